#!/usr/bin/env python
"""Comparator only (never linked into the library): the reference's UN-FUSED step written with torch ops, i.e. cuFFT for
the 7 transforms of `SimulationObject::update()` (simulation_object.rs:475-661) and one element-wise kernel per array
operation, the way ArrayFire executes it -- and the bare cuFFT transforms.  One stream of the bench workload at a time
(streams are independent), fp64, 512^3 by default.

    python scripts/cufft_comparator.py [size] [steps]      -> one JSON line
"""
import json
import math
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
L, hbar_, M, cfl, T, dumps = 30.0, 0.02, 1e10, 0.02, 400.0, 200          # examples/gaussian-overdensity-mft.toml
POIS = 4.0 * math.pi * 4.49e-12
dx = L / n
k1 = torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=dev)
k2 = ((k1[:, None, None] ** 2 + k1[None, :, None] ** 2) + k1[None, None, :] ** 2) * (2 * math.pi) ** 2   # spec_grid
k2max = float(k2.max())
x = (torch.arange(n, dtype=torch.float64, device=dev) * 2 + 1) * dx / 2
g = torch.exp(-0.5 * ((x - 15.0) / 10.0) ** 2)
psi = (g[:, None, None] * g[None, :, None] * g[None, None, :]).to(torch.complex128)
psi = psi * math.sqrt(dx ** -3 / float((psi.abs() ** 2).sum()))
psi = psi + (torch.randn_like(psi.real) + 1j * torch.randn_like(psi.real)) / (2e5 * math.sqrt(dx ** 3))


def potential(psi):
    rho = (M * (psi * psi.conj()).real).to(torch.complex128)               # calculate_density: rho stored complex
    rk = torch.fft.fftn(rho, norm="ortho")
    pk = (-POIS) * rk / k2
    pk[0, 0, 0] = 0.0                                                       # NaN -> 0
    return torch.fft.ifftn(pk, norm="ortho")


def update(psi, psik, t, cur):
    phi = potential(psi)
    pmax = float(phi.abs().max())                                           # max_all: host sync
    kin = cfl * 2 * L / math.sqrt(k2max) / hbar_
    pot = cfl * 2 * math.pi * hbar_ / (2 * pmax)
    nxt = (cur + 1) * T / dumps - t
    dt = min(kin, pot, nxt)
    kev = torch.exp(-1j * (dt / 4 * hbar_) * k2)
    psik = psik * kev
    psi = torch.fft.ifftn(psik, norm="ortho")
    phi = potential(psi)
    psi = psi * torch.exp(-1j * (dt / hbar_) * phi)
    psik = torch.fft.fftn(psi, norm="ortho")
    psik = psik * kev
    psi = torch.fft.ifftn(psik, norm="ortho")
    alias = float(((psik * psik.conj()).real * (k2 > 0.95 * k2max)).sum()) * dx ** 3      # sum_all: host sync
    return psi, psik, t + dt, alias


psik = torch.fft.fftn(psi, norm="ortho")
t = 0.0
psi, psik, t, _ = update(psi, psik, t, 0)                                   # warm-up (cuFFT plans)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    psi, psik, t, alias = update(psi, psik, t, 0)
b.record()
torch.cuda.synchronize()
step_ms = a.elapsed_time(b) / steps
# bare transforms
a.record()
for _ in range(5):
    psik = torch.fft.fftn(psi, norm="ortho")
    psi = torch.fft.ifftn(psik, norm="ortho")
b.record()
torch.cuda.synchronize()
fft_ms = a.elapsed_time(b) / 10
cells = n ** 3
print(json.dumps({"comparator": "torch ops + cuFFT, un-fused reference sequence (7 transforms, ~25 element-wise kernels, 2 host syncs)",
                  "size": n, "steps": steps, "ms_per_step_one_stream": step_ms, "cell_updates_per_s": cells / (step_ms * 1e-3),
                  "cufft_c2c_3d_ms": fft_ms, "cufft_GBps_in_the_3_pass_model": 96.0 * cells / (fft_ms * 1e-3) / 1e9,
                  "cufft_GBps_at_one_read_one_write": 32.0 * cells / (fft_ms * 1e-3) / 1e9,
                  "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30, "alias": alias, "t": t}))
