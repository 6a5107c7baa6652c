#!/usr/bin/env python
"""A/B timing of variant libraries (scripts/build_variant.sh): runs the short 16-stream bench once per library and prints
the step rate and the per-kernel table side by side.   python scripts/ab_bench.py base pi1 pi2 ... [--streams 16]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = [a for a in sys.argv[1:] if not a.startswith("--")]
extra = [a for a in sys.argv[1:] if a.startswith("--")]
streams = "16"
for a in extra:
    if a.startswith("--streams="):
        streams = a.split("=")[1]
coupling = "summed" if "--summed" in extra else "independent"
rows = {}
for nm in names:
    env = dict(os.environ)
    if nm != "base":
        env["MSM_B200_LIB"] = os.path.join(ROOT, "msm_b200", f"libmsm_b200_{nm}.so")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--streams", streams, "--steps", "6", "--warmup", "2",
                          "--no-e2e", "--no-cpu", "--no-summed", "--coupling", coupling], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(nm, "FAILED", out.stderr[-600:])
        continue
    rows[nm] = d
    print(f"{nm:8s} {d['value'] / 1e9:7.3f} G  {d['ms_per_step']:8.2f} ms/step  sm {d['clocks']['sm_mhz']}", flush=True)
kn = []
for d in rows.values():
    for k in d["roofline"]["kernels"]:
        if k["name"] not in kn:
            kn.append(k["name"])
print(f"{'kernel':52s}" + "".join(f"{nm:>10s}" for nm in rows))
for k in kn:
    line = f"{k:52s}"
    for d in rows.values():
        ms = [x["ms"] for x in d["roofline"]["kernels"] if x["name"] == k]
        line += f"{ms[0]:10.1f}" if ms else f"{'-':>10s}"
    print(line)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_" + "_".join(names) + ".json"), "w"))
