"""debug helper (not part of the test-suite): per-step alias mass / psi error on a small blocked-layout grid"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import msm_b200 as m
from oracle import msm_oracle as o
from golden_util import initial_wavefunction, oracle_streams, to_msm_params
size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ps = oracle_streams("spherical-tophat", size, limit=3)
sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=len(ps))
refs = []
for i, p in enumerate(ps):
    psi0 = initial_wavefunction(p)
    sim.set_psi(i, psi0)
    refs.append(o.SimulationObject(p, psi0))
for k in range(4):
    sim.update()
    for i, r in enumerate(refs):
        r.update()
        st = sim.state(i)
        e = np.linalg.norm((sim.get_psi(i) - r.psi).ravel()) / np.linalg.norm(r.psi.ravel())
        ek = np.linalg.norm((sim.grid.get_psik(i) - r.psik).ravel()) / np.linalg.norm(r.psik.ravel())
        print(k, i, "alias", st.alias_mass, r.last_alias_mass, "pmax rel", abs(st.potential_max - r.last_potential_max) / r.last_potential_max,
              "psi err", e, "psik err", ek)
