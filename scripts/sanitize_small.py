"""Small run that touches every kernel variant (for compute-sanitizer): 3-D fused step, 2-D, 1-D, summed mode,
odd stream counts, dumps, potential, device ICs, n = 8 / 16 / 64 / 128."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import msm_b200 as m
from oracle import msm_oracle as o
import golden_util as gu

def rel(a, b): return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())

def run(name, size, ns, steps, coupling=m.COUPLING_INDEPENDENT, dims=None, chunk=0):
    t = gu.load_toml(name, size)
    if dims: t.dims = dims
    ps = list(o.simulation_iter(t))[:ns]
    sim = m.SimulationObject(gu.to_msm_params(ps[0]), n_streams=ns, coupling=coupling, chunk_streams=chunk)
    refs = []
    psi0s = [gu.initial_wavefunction(p) for p in ps]
    for i, a in enumerate(psi0s): sim.set_psi(i, a)
    if coupling == m.COUPLING_SUMMED:
        ens = o.SummedEnsemble(ps[0], psi0s)
        for _ in range(steps): sim.update(); ens.update()
        err = max(rel(sim.get_psi(i), ens.streams[i].psi) for i in range(ns))
    else:
        refs = [o.SimulationObject(p, a) for p, a in zip(ps, psi0s)]
        for _ in range(steps):
            sim.update()
            for r in refs: r.update()
        err = max(rel(sim.get_psi(i), r.psi) for i, r in enumerate(refs))
    sim.grid.get_potential(0); sim.grid.get_psik(0); sim.grid.get_psi_planes(0)
    sim.close()
    print(f"{name} n={size or 16} dims={dims or 3} S={ns} coupling={coupling}: psi rel-L2 {err:.2e}", flush=True)
    assert err < 1e-10

run("spherical-tophat", None, 3, 2)
run("spherical-tophat", 8, 2, 2)
run("spherical-tophat", 32, 5, 2, chunk=2)
run("spherical-tophat-cosmo", None, 2, 2)
run("spherical-tophat", 32, 2, 2, dims=2)
run("repro-planeWave1d", None, 2, 2)
run("spherical-tophat", None, 3, 2, coupling=m.COUPLING_SUMMED, chunk=2)
run("spherical-tophat", 32, 3, 2, coupling=m.COUPLING_SUMMED, chunk=2)     # real-field solve: n/2 = 16 point R2C / C2R
run("spherical-tophat", 64, 2, 1, coupling=m.COUPLING_SUMMED)
run("spherical-tophat", 64, 2, 1)
run("spherical-tophat", 128, 2, 1)
os.environ["MSM_B200_LB"] = "2"
run("spherical-tophat", None, 3, 2)
os.environ["MSM_B200_LB"] = "0"; os.environ["MSM_B200_FUSE"] = "0"
run("spherical-tophat", None, 3, 2)
a = np.random.default_rng(0).standard_normal((2, 256, 256)) + 0j
assert rel(m.inverse(m.forward(a, 2), 2), a) < 1e-13
print("sanitize_small ok")
