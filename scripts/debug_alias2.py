"""debug helper: where (i, j, k) the fused step differs from the oracle after one update"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import msm_b200 as m
from oracle import msm_oracle as o
from golden_util import initial_wavefunction, oracle_streams, to_msm_params
size = 64
ps = oracle_streams("spherical-tophat", size, limit=3)
sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=len(ps))
refs = []
for i, p in enumerate(ps):
    psi0 = initial_wavefunction(p)
    sim.set_psi(i, psi0)
    refs.append(o.SimulationObject(p, psi0))
sim.update()
for i, r in enumerate(refs):
    r.update()
    d = np.abs(sim.grid.get_psik(i) - r.psik)
    scale = np.abs(r.psik).max()
    bad = np.argwhere(d > 1e-9 * scale)
    print("stream", i, "bad", len(bad), "of", d.size, "max rel", d.max() / scale)
    if len(bad):
        for ax, nm in enumerate("ijk"):
            u, c = np.unique(bad[:, ax], return_counts=True)
            print("   axis", nm, "values", u[:24], "n", len(u), "counts", c[:12])
        print("   k//8 %4 histogram", np.bincount((bad[:, 2] // 8) % 4, minlength=4), " i histogram /8", np.bincount(bad[:, 0] // 8, minlength=8))
        b = tuple(bad[0]); print("   first bad", b, sim.grid.get_psik(i)[b], r.psik[b])
