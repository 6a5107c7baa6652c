#!/usr/bin/env python
"""Static check for the store / back-edge hazard of DESIGN.md section 6 in compiled pass kernels.

An instruction that reads its source registers late (global / shared stores and loads) sets a READ BARRIER; whoever
overwrites one of those registers must first wait for that barrier.  ptxas (CUDA 12.9, sm_100a) was seen to omit the wait
on a path through a loop back-edge.  For every backward branch B -> T of every kernel this script collects the
instructions before B whose read barrier is still open at B, then explores the code reachable from T (both directions of
conditional branches, up to a depth) and reports any instruction that overwrites one of the registers still owed to
such an instruction before the path has waited for its barrier (or passed a MEMBAR, which completes only after every
earlier memory instruction).

    python scripts/check_war_hazard.py msm_b200/csrc/build/fft_512.o [more objects]      exit status 1 if anything is found
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_ctrl import all_kernels  # noqa: E402

NO_DEST = ("ST", "BRA", "BAR", "EXIT", "ISETP", "DSETP", "FSETP", "MEMBAR", "CCTL", "NOP", "BSYNC", "BSSY", "RED", "PLOP",
           "WARPSYNC", "ERRBAR", "UISETP", "UIADD", "UIMAD", "ULEA", "USHF", "UMOV", "USEL", "ULOP", "UPLOP", "LDCU", "S2UR",
           "R2UR", "UFLO", "UPOPC", "UBREV", "UBMSK", "VOTEU", "UP2UR", "UR2UP", "UCLEA", "UF2FP", "CALL", "RET", "YIELD")
LOOKBACK, DEPTH = 4000, 1500


def strip_pred(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text)


def reg_range(tok, n):
    m = re.match(r"R(\d+)$", tok.split(".")[0])
    if not m:
        return set()
    b = int(m.group(1))
    return set(range(b, b + n))


def width(op):
    """registers written by the destination operand"""
    if ".128" in op:
        return 4
    if op[0] == "D" or ".64" in op or ".WIDE" in op or ".F64" in op or (op.startswith("CS2R") and ".32" not in op):
        return 2
    return 1


def dest_regs(text):
    t = strip_pred(text)
    op, _, rest = t.partition(" ")
    if op.startswith(NO_DEST):
        return set()
    first = rest.split(",")[0].strip()
    if first.startswith("P") or first == "PT":            # e.g. SHFL.BFLY PT, R3, ... / IADD3 R0, P1, ...
        parts = [x.strip() for x in rest.split(",")]
        first = parts[1] if len(parts) > 1 else ""
    return reg_range(first, width(op))


def source_regs(text):
    t = strip_pred(text)
    op, _, rest = t.partition(" ")
    regs = set()
    toks = re.findall(r"(?<![A-Za-z])R(\d+)(\.64)?", rest)      # (not the R of a uniform register: desc[UR12])
    is_store = op.startswith("ST")
    skip_first = not (is_store or op.startswith(NO_DEST))
    for k, (r, wide) in enumerate(toks):
        if skip_first and k == 0:
            continue
        n = 2 if wide else 1
        regs |= set(range(int(r), int(r) + n))
    if is_store:                                             # the data operand is as wide as the access
        m = re.search(r"\],\s*R(\d+)", rest)
        if m:
            regs |= set(range(int(m.group(1)), int(m.group(1)) + (4 if ".128" in op else 2 if ".64" in op else 1)))
    return regs


def branch_target(text):
    m = re.search(r"\bBRA(?:\.U|\.DIV|\.CONV)*\s+(?:!?U?P\d+,\s*)?(?:P\d,\s*)?0x([0-9a-f]+)", strip_pred(text))
    return format(int(m.group(1), 16) & 0xFFFFF, "04x") if m else None


def check(obj):
    found = []
    for name, ins in all_kernels(obj).items():
        index = {x["addr"]: i for i, x in enumerate(ins)}
        for i, x in enumerate(ins):
            tgt = branch_target(x["text"])
            if tgt is None or tgt not in index or index[tgt] > i:
                continue
            pending = {}
            for y in ins[max(0, i - LOOKBACK):i + 1]:
                for b in range(6):
                    if y["wait"] >> b & 1:
                        pending.pop(b, None)
                if "MEMBAR" in y["text"]:
                    pending.clear()
                if y["rd"] != 7:
                    pending.setdefault(y["rd"], set()).update(source_regs(y["text"]))
            if not pending:
                continue
            seen = set()
            work = [(index[tgt], {b: frozenset(r) for b, r in pending.items()}, 0)]
            while work:
                j, pend, depth = work.pop()
                while j < len(ins) and depth < DEPTH and pend:
                    key = (j, tuple(sorted(pend)))
                    if key in seen:
                        break
                    seen.add(key)
                    y = ins[j]
                    pend = {b: r for b, r in pend.items() if not (y["wait"] >> b & 1)}
                    if "MEMBAR" in y["text"]:
                        pend = {}
                    d = dest_regs(y["text"])
                    for b, rs in pend.items():
                        if d & rs:
                            found.append(f"{name}: back-edge {x['addr']}->{tgt}: {y['addr']} {y['text'][:44]} overwrites "
                                         f"R{sorted(d & rs)} still owed to an instruction with read barrier {b}")
                    if y["rd"] != 7:                        # the barrier is re-armed by a later instruction: stop tracking it
                        pend.pop(y["rd"], None)
                    t2 = branch_target(y["text"])
                    op = strip_pred(y["text"]).split(" ")[0]
                    if t2 is not None and t2 in index and index[t2] > j:
                        work.append((index[t2], dict(pend), depth + 1))
                        if not y["text"].startswith("@") and "BRA.U" not in y["text"]:
                            break                            # unconditional forward branch: no fall-through
                    elif op.startswith(("EXIT", "RET")) and not y["text"].startswith("@"):
                        break
                    elif t2 is not None:                     # another backward branch: leave it to its own analysis
                        if not y["text"].startswith("@") and "BRA.U" not in y["text"]:
                            break
                    j += 1
                    depth += 1
    return sorted(set(found))


if __name__ == "__main__":
    total = 0
    for o in sys.argv[1:]:
        for line in check(o):
            print(line[:220])
            total += 1
    print("hazards found:", total)
    sys.exit(1 if total else 0)
