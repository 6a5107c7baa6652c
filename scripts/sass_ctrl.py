#!/usr/bin/env python
"""Print the scheduling control fields of every SASS instruction of one kernel (sm_70+ encoding: bits 105-125 of the
128-bit instruction = stall count, yield, write-barrier index, read-barrier index, wait mask, reuse).

    python scripts/sass_ctrl.py <object or cubin> <substring of the mangled kernel name> [first_addr last_addr]

Used to document the store / prologue hazard of DESIGN.md section 6: build fft_inst.cu with -DMSM_FFT_N=64
-DMSM_NO_TILE_FENCE and look at the last STG.E.128 of the tile body (read barrier 2) and the first instructions at the
tile-loop top that overwrite their data registers without waiting for that barrier."""
import re
import subprocess
import sys


def all_kernels(obj):
    """{mangled name: [instruction dicts]} for every kernel of an object file / cubin (one cuobjdump call)"""
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    kernels, cur, i = {}, None, 0
    while i < len(text):
        line = text[i]
        f = re.search(r"Function : (\S+)", line)
        if f:
            cur = kernels.setdefault(f.group(1), [])
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line) if cur is not None else None
        if m and i + 1 < len(text):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", text[i + 1])
            if m2:
                ctrl = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
                cur.append(dict(addr=m.group(1), text=m.group(2).strip(), stall=ctrl & 0xF, wr=(ctrl >> 5) & 7,
                                rd=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3F))
                i += 2
                continue
        i += 1
    return kernels


def kernel_instructions(obj, key):
    for name, ins in all_kernels(obj).items():
        if key in name:
            return ins
    return []


if __name__ == "__main__":
    ins = kernel_instructions(sys.argv[1], sys.argv[2])
    lo, hi = (sys.argv[3], sys.argv[4]) if len(sys.argv) > 4 else ("0000", "ffff")
    for x in ins:
        if lo <= x["addr"] <= hi:
            b = lambda v: "-" if v == 7 else str(v)
            print(f"{x['addr']}  {x['text'][:64]:64s} stall {x['stall']:2d}  wr {b(x['wr'])}  rd {b(x['rd'])}  wait {x['wait']:06b}")
