#!/usr/bin/env python
"""Golden trajectory from the REFERENCE'S OWN legacy Python integrator, run in this container.

`/root/reference/python_deprecated/gravSolver.py` (`Grav3D.Update`, :91-118, with `compute_phi`, :77-89) is the authors'
earlier implementation of the same split step the Rust `SimulationObject::update()` performs (simulation_object.rs:504-581):
    psi_k *= exp(-i dt hbar k^2 / (4 m));  rho = Mtot |psi|^2;  phi_k = -C rho_k / k^2, phi_k[0] = 0;
    psi *= exp(-i dt m phi / hbar);         psi_k *= exp(-i dt hbar k^2 / (4 m))
with a FIXED dt.  The Rust crate has no trajectory fixtures and cannot be built here, so this is the one piece of reference
code that can be executed to pin the integrator step of oracle/msm_oracle.py (and through it the CUDA path).

Its only missing dependency is `pyfftw`: a 15-line stand-in (same call signature, numpy.fft / pocketfft underneath,
inverse normalised by 1/N^3 as pyfftw.FFTW.__call__ does by default) is injected before the import.  Nothing else is
replaced: the class, its ColdGauss initial condition, Update and compute_phi run unmodified from /root/reference.

    python tests/golden/make_legacy_golden.py     ->  tests/golden/legacy_grav3d_traj.npz
Physical scalars: those of examples/spherical-tophat.toml (L = 30, hbar_ = 0.05, M = 1e11, C = 4 pi 4.49e-12), n = 16,
dt = 0.2 (that example's dump-limited step), Gaussian of width 3 in the box centre; psi after 1, 10 and 40 steps.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/python_deprecated"

fake = types.ModuleType("pyfftw")


class FFTW:   # pyfftw.FFTW(input, output, axes=..., direction=...): calling it transforms input into output
    def __init__(self, input_array, output_array, axes=(-1,), direction="FFTW_FORWARD", **kw):
        self.i, self.o, self.axes, self.fwd = input_array, output_array, tuple(axes), direction == "FFTW_FORWARD"

    def __call__(self):
        r = np.fft.fftn(self.i, axes=self.axes) if self.fwd else np.fft.ifftn(self.i, axes=self.axes)
        self.o[...] = r
        return self.o


fake.FFTW = FFTW
fake.empty_aligned = lambda shape, dtype="complex128", **kw: np.empty(shape, dtype=dtype)
sys.modules["pyfftw"] = fake
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import gravSolver  # noqa: E402  (the reference's file, unmodified)
import scipy.fftpack as sp  # noqa: E402

N, L = 16, 30.0
HBAR_, MTOT, C = 0.05, 1e11, 4.0 * np.pi * 4.49e-12
DT = 0.2
dx = L / N

s = types.SimpleNamespace()
s.hbar, s.mpart = HBAR_, 1.0                  # only hbar / mpart = hbar_ enters (gravSolver.py:99,109)
s.Mtot, s.C = MTOT, C
s.kx = 2 * np.pi * sp.fftfreq(N, d=dx)        # SimObj.py:70-76 MakeSpecGrid
ky, kx, kz = np.meshgrid(s.kx, s.kx, s.kx)
s.spec_grid = kx ** 2 + ky ** 2 + kz ** 2
s.k_cutoff_frac, s.P_thresh, s.cf, s.dt = 0.95, 0.01, 0.1, DT   # CheckAlias bookkeeping (does not touch psi)

g = gravSolver.Grav3D(N, 1)
psi0 = np.array(g.ColdGauss([0.0, 0.0, 0.0], [3.0, 3.0, 3.0], N, L, dtype_="complex128"), dtype=np.complex128)
assert abs(np.sum(np.abs(psi0) ** 2) * dx ** 3 - 1.0) < 1e-12
g.psi = psi0.copy()
g.fft_out = np.empty((N, N, N), dtype=np.complex128)
g.fft_out_phi = np.empty((N, N, N), dtype=np.complex128)

out = {"psi0": psi0}
with np.errstate(divide="ignore", invalid="ignore"):
    for steps in (1, 10, 40):
        # `while T < Tgoal: dt_ = min(dt, Tgoal - T)` (gravSolver.py:93-94): `steps` steps of dt (when the accumulated T
        # falls an ulp short of Tgoal the loop adds one step of ~1e-16, which moves psi by nothing)
        g.Update(DT, steps * DT, s)
        assert abs(g.T - steps * DT) < 1e-12, g.T
        out[f"psi_{steps:03d}"] = np.array(g.psi, dtype=np.complex128)
rho = MTOT * np.abs(psi0) ** 2
with np.errstate(divide="ignore", invalid="ignore"):
    out["phi0"] = np.array(g.compute_phi(rho + 0j, s).real)
out["params"] = np.array([N, L, HBAR_, MTOT, C, DT])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "legacy_grav3d_traj.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
