"""Size-independent properties of the CUDA path at BASELINE.json's full grid sizes (256^3, 512^3), where the CPU
oracle is too slow to run whole trajectories: unitarity, Parseval, linearity, exact time reversal, zero-step
identity, stream independence -- plus one direct oracle comparison at 256^3."""
import numpy as np
import pytest

import msm_b200 as m
from oracle import msm_oracle as o
from conftest import rel_l2
from golden_util import oracle_streams, to_msm_params

pytestmark = pytest.mark.gpu


def gaussian_params(size):
    return oracle_streams("gaussian-overdensity-mft", size, limit=1)[0]


def test_fft_512_round_trip_parseval():
    rng = np.random.default_rng(1)
    a = np.empty((512, 512, 512), dtype=np.complex128)
    for i in range(512):                               # plane by plane keeps host temporaries small
        a[i] = rng.standard_normal((512, 512)) + 1j * rng.standard_normal((512, 512))
    f = m.forward(a)
    e0, e1 = np.vdot(a, a).real, np.vdot(f, f).real
    assert abs(e1 - e0) < 1e-12 * e0                   # unitary (simulator/tests/fft.rs:44-50)
    assert abs(f[0, 0, 0] - a.sum() / 512 ** 1.5) < 1e-9 * abs(f[0, 0, 0])
    back = m.inverse(f)
    assert rel_l2(back, a) < 1e-13


def test_fft_linearity_256():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((256,) * 3) + 1j * rng.standard_normal((256,) * 3)
    b = rng.standard_normal((256,) * 3) + 1j * rng.standard_normal((256,) * 3)
    fa, fb, fab = m.forward(a), m.forward(b), m.forward(a + (2.0 - 0.5j) * b)
    assert rel_l2(fab, fa + (2.0 - 0.5j) * fb) < 1e-13


@pytest.mark.parametrize("size", [256, 512])
def test_step_properties_at_full_size(size):
    p = gaussian_params(size)
    sim = o.SimulationObject(p, np.zeros((2, 2, 2), dtype=np.complex128))
    ctx = m.Context(3, size, 2, p.dx, sim.density_prefactor(), sim.poisson_coeff(), p.k2_cutoff, chunk_streams=2)
    ctx.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    ctx.ic_copy(1, 0)
    ctx.sample_perturbation(0, "Wigner", 11, 1e10)
    ctx.sample_perturbation(1, "Wigner", 12, 1e10)
    dv = p.dx ** 3
    psik0 = ctx.get_psik(0)
    n0 = np.vdot(psik0, psik0).real * dv
    assert abs(n0 - 1.0) < 2e-2                         # normalised IC + Wigner noise (cells / (2 n_tot) of extra norm)
    # zero-dt step is the identity on psi_k
    ctx.step(np.zeros(2), np.zeros(2))
    assert rel_l2(ctx.get_psik(0), psik0) < 1e-14
    # a real step: norm conserved (every operator is unitary), alias mass tiny for a smooth field + noise
    pm = ctx.potential_max()
    dt = p.cfl * np.pi * p.hbar_ / pm                   # simulation_object.rs:906-909
    alias = ctx.step(dt * p.hbar_ / 4.0, dt / p.hbar_)
    k1 = ctx.get_psik(0)
    assert abs(np.vdot(k1, k1).real * dv - n0) < 1e-12
    assert rel_l2(k1, psik0) > 1e-6                     # it did move
    assert alias[0] < 1e-5 and alias[1] < 1e-5          # only the white Wigner noise reaches beyond the cutoff
    # exact time reversal of the split step: D(-dt/2) K(-dt) D(-dt/2) undoes D(dt/2) K(dt) D(dt/2)
    ctx.step(-dt * p.hbar_ / 4.0, -dt / p.hbar_)
    assert rel_l2(ctx.get_psik(0), psik0) < 1e-12
    ctx.close()


@pytest.mark.parametrize("size,streams", [(512, 3), (256, 5), (64, 3), (1024, 0)])
def test_fused_step_equals_the_plain_pass_sequence(monkeypatch, size, streams):
    """The fused kernels (two transforms per tile, several tiles per CTA) against the un-fused 3+3 pass sequence of the
    same library (MSM_B200_FUSE=0), whose plain passes are pinned to pocketfft by the FFT tests: two real steps, odd
    stream count (a group with a single stream).  Caught a store/prologue hazard at tile boundaries (fft_pass.cuh)."""
    if size == 1024:                                    # 2-D: the only shape at which 1024-point lines are affordable
        rng = np.random.default_rng(3)
        a = rng.standard_normal((5, 1024, 1024)) + 1j * rng.standard_normal((5, 1024, 1024))
        import scipy.fft as sf
        assert rel_l2(m.forward(a, 2), sf.fftn(a, axes=(1, 2), norm="ortho")) < 1e-13
        return
    p = gaussian_params(size)
    osim = o.SimulationObject(p, np.zeros((2, 2, 2), dtype=np.complex128))
    results = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("MSM_B200_FUSE", fuse)
        ctx = m.Context(3, size, streams, p.dx, osim.density_prefactor(), osim.poisson_coeff(), p.k2_cutoff, chunk_streams=4)
        ctx.ic_cold_gauss(0, [15.0] * 3, [6.0 + streams] * 3)
        for s in range(1, streams):
            ctx.ic_copy(s, 0)
        for s in range(streams):
            ctx.sample_perturbation(s, "Wigner", 21 + s, 1e6)
        out = []
        for step in range(2):
            pm = ctx.potential_max()
            dt = p.cfl * np.pi * p.hbar_ / pm * (1.0 + 0.1 * np.arange(streams))     # a different dt per stream
            alias = ctx.step(dt * p.hbar_ / 4.0, dt / p.hbar_)
            out.append((pm, alias))
        out.append([ctx.get_psik(s) for s in (0, streams - 1)])
        results.append(out)
        ctx.close()
    fused, plain = results
    for step in range(2):
        assert np.all(np.abs(fused[step][0] - plain[step][0]) <= 1e-12 * plain[step][0])         # max|phi|
        assert np.all(np.abs(fused[step][1] - plain[step][1]) <= 1e-10 * plain[step][1] + 1e-30)  # alias mass
    for a, b in zip(fused[2], plain[2]):
        assert rel_l2(a, b) < 1e-12


def test_streams_do_not_leak_into_each_other_256():
    """Two streams share one complex pair buffer for rho/phi; a stream's result must not depend on its partner."""
    p = gaussian_params(256)
    sim = o.SimulationObject(p, np.zeros((2, 2, 2), dtype=np.complex128))
    args = (3, 256, 2, p.dx, sim.density_prefactor(), sim.poisson_coeff(), p.k2_cutoff)
    a = m.Context(*args, chunk_streams=2)
    a.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    a.ic_cold_gauss(1, [12.0, 17.0, 15.0], [4.0, 5.0, 6.0])       # a very different partner
    b = m.Context(*args, chunk_streams=2)
    b.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    b.ic_cold_gauss(1, [15.0] * 3, [10.0] * 3)
    pa, pb = a.potential_max(), b.potential_max()
    assert abs(pa[0] - pb[0]) <= 1e-13 * pb[0] and pa[1] > 2 * pa[0]
    dt = np.full(2, 0.3)
    a.step(dt * p.hbar_ / 4.0, dt / p.hbar_)
    b.step(dt * p.hbar_ / 4.0, dt / p.hbar_)
    assert rel_l2(a.get_psik(0), b.get_psik(0)) < 1e-13
    a.close()
    b.close()


def test_two_steps_match_oracle_at_256():
    p = gaussian_params(256)
    psi0 = o.cold_gauss([15.0] * 3, [10.0] * 3, p)
    ref = o.SimulationObject(p, psi0)
    sim = m.SimulationObject(to_msm_params(p), n_streams=1)
    sim.set_psi(0, psi0)
    for _ in range(2):
        sim.update()
        ref.update()
        st = sim.state(0)
        assert abs(st.dt - ref.last_dt) <= 1e-13 * ref.last_dt
        assert abs(st.potential_max - ref.last_potential_max) <= 1e-12 * ref.last_potential_max
    assert rel_l2(sim.get_psi(0), ref.psi) < 1e-10
    sim.close()


def test_tma_variant_of_the_plain_pass_is_bit_identical_512(monkeypatch):
    """MSM_B200_TMA=1 routes the plain strided 512-point passes (Poisson y passes, dt-potential y pass) through the
    cp.async.bulk.tensor kernel of fft_tma.cu: same butterflies, same order -> the same bits after two steps."""
    p = gaussian_params(512)
    osim = o.SimulationObject(p, np.zeros((2, 2, 2), dtype=np.complex128))
    out = []
    for tma in ("0", "1"):
        monkeypatch.setenv("MSM_B200_TMA", tma)
        ctx = m.Context(3, 512, 3, p.dx, osim.density_prefactor(), osim.poisson_coeff(), p.k2_cutoff, chunk_streams=4)
        ctx.ic_cold_gauss(0, [15.0] * 3, [9.0] * 3)
        for s in (1, 2):
            ctx.ic_copy(s, 0)
        for s in range(3):
            ctx.sample_perturbation(s, "Wigner", 31 + s, 1e6)
        ctx.profile_enable(True)
        res = []
        for _ in range(2):
            pm = ctx.potential_max()
            dt = p.cfl * np.pi * p.hbar_ / pm
            res.append((pm, ctx.step(dt * p.hbar_ / 4.0, dt / p.hbar_)))
        names = [r["name"] for r in ctx.profile_read()]
        assert any("[tma]" in nm for nm in names) == (tma == "1")
        res.append([ctx.get_psik(s) for s in (0, 2)])
        out.append(res)
        ctx.close()
    a, b = out
    for k in range(2):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1])
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y)
