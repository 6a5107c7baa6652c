"""First-light diagnostics on a B200: FFT of every size/dims vs scipy, then one step vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.fft as sf
import msm_b200 as m
from oracle import msm_oracle as o

def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)

rng = np.random.default_rng(0)
bad = 0
for dims in (1, 2, 3):
    for n in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        if n ** dims > 2 ** 24:
            continue
        batch = 3
        a = rng.standard_normal((batch,) + (n,) * dims) + 1j * rng.standard_normal((batch,) + (n,) * dims)
        try:
            f = m.forward(a, dims)
            r = sf.fftn(a, axes=tuple(range(1, dims + 1)), norm="ortho")
            e1 = rel(f, r)
            b = m.inverse(f, dims)
            e2 = rel(b, a)
            flag = "" if (e1 < 1e-13 and e2 < 1e-13) else "   <-- BAD"
            bad += bool(flag)
            print(f"fft dims={dims} n={n:5d}  fwd {e1:.2e}  roundtrip {e2:.2e}{flag}", flush=True)
        except Exception as ex:
            bad += 1
            print(f"fft dims={dims} n={n}: EXC {ex}", flush=True)

# spec grid
for dims in (1, 2, 3):
    k2 = m.spec_grid(0.25, dims, 4)
    print("spec_grid", dims, np.array_equal(k2, o.spec_grid(0.25, dims, 4)))

# a few steps vs the oracle on the reference's example configs (fixtures under tests/golden)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_util as gu

def one(name, size=None, steps=3, nstreams=3, expanding=None):
    ps = gu.oracle_streams(name, size, expanding, nstreams)
    sim = m.SimulationObject(gu.to_msm_params(ps[0]), n_streams=len(ps))
    osims = []
    for i, p in enumerate(ps):
        psi0 = gu.initial_wavefunction(p)
        sim.set_psi(i, psi0)
        osims.append(o.SimulationObject(p, psi0))
    for k in range(steps):
        sim.update()
        for i, os_ in enumerate(osims):
            os_.update()
            st = sim.state(i)
            psi = sim.get_psi(i)
            print(f"  step {k} stream {i}: dt {st.dt:.6e} vs {os_.last_dt:.6e} | pmax rel {abs(st.potential_max-os_.last_potential_max)/os_.last_potential_max:.1e}"
                  f" | alias {st.alias_mass:.3e} vs {os_.last_alias_mass:.3e} | psi rel {rel(psi, os_.psi):.2e} | t {st.time:.6e} vs {os_.parameters.time:.6e}", flush=True)
    sim.close()

for name, kw in (("spherical-tophat", {}), ("spherical-tophat-cosmo", {}), ("planeWave3d_e10_sym", {}),
                 ("spherical-tophat", {"size": 64, "nstreams": 2}), ("repro-planeWave1d", {"nstreams": 2}),
                 ("gaussian-overdensity-mft", {"size": 32, "nstreams": 1})):
    print(name, kw, flush=True)
    try:
        one(name, **kw)
    except Exception as ex:
        bad += 1
        print("  EXC", repr(ex), flush=True)
print("bad:", bad)
