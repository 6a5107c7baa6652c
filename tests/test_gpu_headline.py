"""Direct oracle parity AT THE HEADLINE GRID SIZE (512^3, BASELINE configs[4]).

At n = 512 the library runs template instances and a device layout that no smaller grid uses (N = 512 plans, blocked
slow axis, the 16-points-per-thread `drift+alias+inv` kernel, pair buffers + a one-stream group), so the pieces pinned
separately on small grids are checked here in combination, against the CPU oracle itself:
  * one 3-D transform, forward and inverse, against pocketfft                                  <= 1e-13 rel-L2
  * 3 streams (a pair + a one-stream group) x 2 update() against o.SimulationObject            psi <= 1e-10, dt <= 1e-13,
                                                                                               max|phi| <= 1e-12, alias 1e-10
  * 2 streams x 1 update() in the summed-density mode against o.SummedEnsemble                 same bounds
Initial conditions: ColdGauss of examples/gaussian-overdensity-mft.toml + Wigner noise (n_tot = 1e10) from the ORACLE's
sampler.  About two minutes of host time (8 oracle stream-updates of ~10 s); needs ~40 GiB of host memory.
Set MSM_B200_SKIP_HEADLINE=1 to leave these out.
"""
import copy
import os

import numpy as np
import pytest
import scipy.fft as sf

import msm_b200 as m
from oracle import msm_oracle as o
from conftest import rel_l2
from golden_util import oracle_streams, to_msm_params

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MSM_B200_SKIP_HEADLINE") == "1", reason="MSM_B200_SKIP_HEADLINE=1")]

N = 512


def host_gib():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2.0 ** 30
    except (ValueError, OSError):
        return 0.0


needs_memory = pytest.mark.skipif(host_gib() < 48.0, reason="needs ~40 GiB of host memory")


@pytest.fixture(scope="module")
def workload():
    """parameters of BASELINE configs[4] and three Wigner-noised ColdGauss wavefunctions from the oracle's sampler"""
    o.set_workers(os.cpu_count() or 1)
    p = oracle_streams("gaussian-overdensity-mft", N, limit=1)[0]
    p.particle_mass = p.total_mass / 1e10                        # n_tot = 1e10
    p.n_tot = 1e10
    base = o.cold_gauss([15.0] * 3, [10.0] * 3, p)
    psi0s = [o.sample_quantum_perturbation(base, p, {"seed": s, "scheme": "Wigner"}) for s in (1, 2, 3)]
    return p, psi0s


def alias_close(a, b):
    return abs(a - b) <= max(1e-10 * abs(b), 1e-30)


@needs_memory
def test_fft_512_cubed_matches_pocketfft():
    rng = np.random.default_rng(7)
    a = np.empty((N, N, N), dtype=np.complex128)
    for i in range(N):                                           # plane by plane keeps host temporaries small
        a[i] = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    w = os.cpu_count() or 1
    got = m.forward(a)
    ref = sf.fftn(a, norm="ortho", workers=w)
    assert rel_l2(got, ref) < 1e-13
    del got, ref
    got = m.inverse(a)
    ref = sf.ifftn(a, norm="ortho", workers=w)
    assert rel_l2(got, ref) < 1e-13


@needs_memory
def test_three_streams_two_updates_match_the_oracle_at_512(workload):
    p, psi0s = workload
    sim = m.SimulationObject(to_msm_params(p), n_streams=3)      # groups: (0, 1) share a pair buffer, (2) is alone
    for i, a in enumerate(psi0s):
        sim.set_psi(i, a)
    states = []
    for _ in range(2):
        sim.update()
        states.append([sim.state(i) for i in range(3)])
    got = [sim.get_psi(i) for i in range(3)]
    sim.close()
    for i in range(3):                                           # one oracle object at a time (~16 GiB each)
        ref = o.SimulationObject(copy.copy(p), psi0s[i])
        for k in range(2):
            ref.update()
            st = states[k][i]
            assert abs(st.dt - ref.last_dt) <= 1e-13 * ref.last_dt, (i, k, st.dt, ref.last_dt)
            assert abs(st.potential_max - ref.last_potential_max) <= 1e-12 * ref.last_potential_max
            assert alias_close(st.alias_mass, ref.last_alias_mass), (st.alias_mass, ref.last_alias_mass)
            assert abs(st.time - ref.parameters.time) <= 1e-13 * ref.parameters.time
            assert st.n_steps == ref.parameters.n_steps and st.current_dumps == ref.parameters.current_dumps
        err = rel_l2(got[i], ref.psi)
        assert err < 1e-10, (i, err)
        got[i] = None
        del ref


@needs_memory
def test_summed_mode_one_update_matches_the_oracle_at_512(workload):
    p, psi0s = workload
    sim = m.SimulationObject(to_msm_params(p), n_streams=2, coupling=m.COUPLING_SUMMED)
    for i in range(2):
        sim.set_psi(i, psi0s[i])
    sim.update()
    st = sim.state(0)
    got = [sim.get_psi(i) for i in range(2)]
    alias = [sim.state(i).alias_mass for i in range(2)]
    sim.close()
    ens = o.SummedEnsemble(copy.copy(p), psi0s[:2])
    ens.update()
    assert abs(st.dt - ens.head.last_dt) <= 1e-13 * ens.head.last_dt
    assert abs(st.potential_max - ens.head.last_potential_max) <= 1e-12 * ens.head.last_potential_max
    for i in range(2):
        assert rel_l2(got[i], ens.streams[i].psi) < 1e-10
        assert alias_close(alias[i], ens.streams[i].last_alias_mass)
