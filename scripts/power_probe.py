"""Does the pass rate decay under sustained load (power cap)?  Repeats potential_max and prints per-call kernel rates."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msm_b200 as m
S = 8
ctx = m.Context(3, 512, S, 30.0 / 512, 1e10, -5.64e-11, 0.95, chunk_streams=8)
ctx.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
for s in range(1, S):
    ctx.ic_copy(s, 0)
ctx.potential_max()
def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    except Exception as e:
        return str(e)
for rep in range(40):
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.potential_max()
    dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    r = [p for p in prof if p["name"].startswith("fft_pass<512,inv,none,none,y")][0]
    x = [p for p in prof if p["name"].startswith("fft_pass<512,fwd,none,none,x")][0]
    print(f"rep {rep:2d} {dt*1e3/3:7.1f} ms/call  inv-y {r['algorithmic_bytes']/r['ms_total']/1e6:7.0f} GB/s  fwd-x {x['algorithmic_bytes']/x['ms_total']/1e6:7.0f} GB/s | smi {smi()}", flush=True)
    if rep == 19:
        print("sleep 5 s"); time.sleep(5)
