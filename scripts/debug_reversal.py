import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import msm_b200 as m
from oracle import msm_oracle as o
import golden_util as gu

def rel(a, b): return np.linalg.norm((a-b).ravel())/np.linalg.norm(b.ravel())
for size in (64, 128, 256):
    p = gu.oracle_streams("gaussian-overdensity-mft", size, limit=1)[0]
    so = o.SimulationObject(p, np.zeros((2,2,2), dtype=np.complex128))
    for noise in (False, True):
        ctx = m.Context(3, size, 2, p.dx, so.density_prefactor(), so.poisson_coeff(), p.k2_cutoff, chunk_streams=2)
        ctx.ic_cold_gauss(0, [15.0]*3, [10.0]*3); ctx.ic_copy(1, 0)
        if noise:
            ctx.sample_perturbation(0, "Wigner", 11, 1e10); ctx.sample_perturbation(1, "Wigner", 12, 1e10)
        psi_host = ctx.get_psi(0)
        k0 = ctx.get_psik(0)
        pm = ctx.potential_max()
        dt = p.cfl*np.pi*p.hbar_/pm
        for label, dc, kc in (("both", dt*p.hbar_/4, dt/p.hbar_), ("drift only", dt*p.hbar_/4, 0*dt), ("kick only", 0*dt, dt/p.hbar_)):
            ctx.step(dc, kc); k1 = ctx.get_psik(0)
            ctx.step(-dc, -kc); k2 = ctx.get_psik(0)
            print(size, "noise" if noise else "smooth", label, "reversal err %.2e"%rel(k2, k0), "moved %.2e"%rel(k1,k0), flush=True)
        if size <= 128 or True:
            # oracle forced-dt step on the same IC
            s = o.SimulationObject(p, psi_host)
            s.psik = o.forward(s.psi)
            kev = np.exp(complex(0, -dt[0]/4*p.hbar_)*p.spec_grid)
            s.psi = o.inverse(s.psik*kev); s.calculate_potential()
            s.psi = s.psi*np.exp(complex(0,-dt[0]/p.hbar_)*s.phi)
            s.psik = o.forward(s.psi)*kev
            ctx.set_psi(0, psi_host); ctx.step(dt*p.hbar_/4, dt/p.hbar_)
            print(size, "noise" if noise else "smooth", "vs oracle psik %.2e"%rel(ctx.get_psik(0), s.psik), flush=True)
        ctx.close()
