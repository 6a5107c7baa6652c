"""Where the end-to-end time of msm_sim_run_streams goes: wall clock vs the CUDA-event time of every kernel launched
inside it (msm_profile_*), for the bench workload.  python scripts/e2e_breakdown.py [streams] [steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import msm_b200 as m
import bench

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
size, cells = 512, 512 ** 3
params = bench.workload_params(size)
sim = m.SimulationObject(params, n_streams=streams, chunk_streams=8)
g = sim.grid
g.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
g.sample_perturbation(0, "Wigner", 1, 1e10)
host = torch.empty(2 * cells, dtype=torch.float64, pin_memory=True)
re = torch.empty(cells, dtype=torch.float64, pin_memory=True)
im = torch.empty(cells, dtype=torch.float64, pin_memory=True)
hnp, renp, imnp = host.numpy(), re.numpy(), im.numpy()
hnp[:] = g.get_psi(0).reshape(-1).view(np.float64)
sim.close()
for label, prof in (("plain", False), ("profiled", True)):
    sim = m.SimulationObject(params, n_streams=streams, chunk_streams=8)
    g = sim.grid
    if prof:
        g.profile_enable(True)
    t0 = time.perf_counter()
    sim.run_streams(list(range(streams)), [hnp] * streams, [renp] * streams, [imnp] * streams, max_updates=steps)
    wall = time.perf_counter() - t0
    print(f"{label}: wall {wall:.3f} s -> {cells * streams * steps / wall / 1e9:.2f} G cell-updates/s")
    if prof:
        rec = g.profile_read()
        tot = sum(r["ms_total"] for r in rec)
        print(f"  kernels on the compute stream: {tot / 1e3:.3f} s in {sum(r['launches'] for r in rec)} launches")
        for r in sorted(rec, key=lambda r: -r["ms_total"]):
            print(f"    {r['name']:46s} {r['launches']:5d} {r['ms_total']:9.1f} ms  {r['algorithmic_bytes'] / max(r['ms_total'], 1e-9) / 1e6:7.1f} GB/s")
    sim.close()
