"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Run on a B200: pytest -m gpu.

Tolerances (fp64, relative L2 unless noted; measured values are ~1e-15):
  one FFT vs pocketfft            <= 1e-13
  phi, psi, psi_k after one step  <= 1e-12
  dt / dtau                       <= 1e-12 relative (1e-13 on the golden trajectories); time, tau <= 1e-13
  max|phi|                        <= 1e-12 relative
  alias mass                      <= 1e-10 relative or 1e-30 absolute (it is round-off noise when nothing aliases)
  trajectories (up to 200 steps)  <= 1e-10   (the north-star bound)
"""
import copy
import os

import numpy as np
import pytest
import scipy.fft as sf

import msm_b200 as m
from msm_b200 import _lib
from oracle import msm_oracle as o
from conftest import rel_l2
from golden_util import GOLDEN, initial_wavefunction, oracle_streams, to_msm_params

pytestmark = pytest.mark.gpu


def alias_close(a, b):
    return abs(a - b) <= max(1e-10 * abs(b), 1e-30)


# ---------------------------------------------------------------------------------------------------------------
# FFT layer (utils/fft.rs) -- rows a3, a4
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [1, 2, 3])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024])
def test_fft_matches_pocketfft(dims, n):
    if n ** dims > 2 ** 24:
        pytest.skip("covered by the large-grid property tests")
    rng = np.random.default_rng(n * 10 + dims)
    batch = 3
    a = rng.standard_normal((batch,) + (n,) * dims) + 1j * rng.standard_normal((batch,) + (n,) * dims)
    f = m.forward(a, dims)
    ref = sf.fftn(a, axes=tuple(range(1, dims + 1)), norm="ortho")
    assert rel_l2(f, ref) < 1e-13
    assert rel_l2(m.inverse(f, dims), a) < 1e-13
    ib = m.inverse(a, dims)
    assert rel_l2(ib, sf.ifftn(a, axes=tuple(range(1, dims + 1)), norm="ortho")) < 1e-13


@pytest.mark.parametrize("K,S,value,length", [(1, 8, complex(0.0, 128.0 ** -0.5), 128.0), (1, 2, 2j, None),
                                              (2, 2, 2j, None), (3, 2, 2j, None), (2, 8, 2j, None), (3, 8, 2j, None)])
def test_reference_fft_round_trips(K, S, value, length):
    # simulator/tests/fft.rs:2-601
    values = np.full((S,) * K, value, dtype=np.complex128)
    fk = m.forward(values)
    if length is not None:
        assert o.check_norm(fk, length / S, K)
    assert abs(np.sum(np.abs(fk) ** 2) - np.sum(np.abs(values) ** 2)) < 1e-12 * values.size
    assert np.sum(np.abs(m.inverse(fk) - values)) < 1e-12


def test_fft_ragged_batches():
    rng = np.random.default_rng(5)
    for batch in (1, 2, 5, 9, 17):          # odd groups, more than one chunk
        a = rng.standard_normal((batch, 16, 16, 16)) + 1j * rng.standard_normal((batch, 16, 16, 16))
        assert rel_l2(m.forward(a, 3), sf.fftn(a, axes=(1, 2, 3), norm="ortho")) < 1e-13


def test_spec_grid_bit_exact():
    # utils/fft.rs:185-214 demands equality, not closeness
    for dims in (1, 2, 3):
        for n, dx in ((4, 0.25), (16, 30.0 / 16), (64, 0.1437)):
            assert np.array_equal(m.spec_grid(dx, dims, n), o.spec_grid(dx, dims, n))


# ---------------------------------------------------------------------------------------------------------------
# grid level: potential, one step  -- rows a5, a6, a9, a10, a12
# ---------------------------------------------------------------------------------------------------------------
def make_ctx(p, n_streams, **kw):
    sim = o.SimulationObject(p, np.zeros((p.size,) * p.dims, dtype=np.complex128))
    return m.Context(p.dims, p.size, n_streams, p.dx, sim.density_prefactor(), sim.poisson_coeff(), p.k2_cutoff, **kw)


@pytest.mark.parametrize("name,size", [("spherical-tophat", None), ("spherical-tophat", 32),
                                       ("spherical-tophat-cosmo", None), ("repro-planeWave1d", None)])
def test_potential_matches_oracle(name, size):
    ps = oracle_streams(name, size, limit=3)
    ctx = make_ctx(ps[0], len(ps))
    sims = []
    for i, p in enumerate(ps):
        psi0 = initial_wavefunction(p)
        ctx.set_psi(i, psi0)
        s = o.SimulationObject(p, psi0)
        s.calculate_potential()
        sims.append(s)
    pm = ctx.potential_max()
    for i, s in enumerate(sims):
        want = np.max(np.abs(s.phi))
        assert abs(pm[i] - want) <= 1e-12 * want
        phi = ctx.get_potential(i)
        assert rel_l2(phi, s.phi.real) < 1e-12
    ctx.close()


def test_single_step_fields_match_oracle():
    ps = oracle_streams("spherical-tophat", limit=4)
    ctx = make_ctx(ps[0], len(ps))
    sims = []
    for i, p in enumerate(ps):
        psi0 = initial_wavefunction(p)
        ctx.set_psi(i, psi0)
        sims.append(o.SimulationObject(p, psi0))
    dts = np.array([0.05, 0.11, 0.2, 0.31])       # different dt per stream
    hb = ps[0].hbar_
    for i, s in enumerate(sims):                  # the oracle's step with a forced dt
        s.psik = o.forward(s.psi)
        kev = np.exp(complex(0, -dts[i] / 4.0 * hb) * s.parameters.spec_grid)
        s.psik = s.psik * kev
        s.psi = o.inverse(s.psik)
        s.calculate_potential()
        s.psi = s.psi * np.exp(complex(0, -dts[i] / hb) * s.phi)
        s.psik = o.forward(s.psi) * kev
        s.psi = o.inverse(s.psik)
        s.check_alias()
    alias = ctx.step(dts * hb / 4.0, dts / hb)
    for i, s in enumerate(sims):
        assert rel_l2(ctx.get_psik(i), s.psik) < 1e-12
        assert rel_l2(ctx.get_psi(i), s.psi) < 1e-12
        re, im = ctx.get_psi_planes(i)
        assert np.array_equal(re + 1j * im, ctx.get_psi(i))
    # pipelined dump of several streams (row f-2), in a scrambled order
    order = [2, 0, 3, 1, 2]
    res = [np.empty((16,) * 3) for _ in order]
    ims = [np.empty((16,) * 3) for _ in order]
    ctx.get_psi_many(order, res, ims)
    for j, i in enumerate(order):
        assert np.array_equal(res[j] + 1j * ims[j], ctx.get_psi(i))
    for i, s in enumerate(sims):
        assert alias_close(alias[i], s.last_alias_mass)
    ctx.close()


def test_alias_mass_when_power_sits_above_the_cutoff():
    """A wavefunction with real power beyond k2_cutoff: the masked sum must match to 1e-10 (not just be ~0)."""
    p = oracle_streams("spherical-tophat", limit=1)[0]
    p.k2_cutoff = 0.3
    p.__post_init__()
    rng = np.random.default_rng(3)
    psi0 = o.normalize(rng.standard_normal((16,) * 3) + 1j * rng.standard_normal((16,) * 3), p.dx, 3)
    ctx = make_ctx(p, 1)
    ctx.set_psi(0, psi0)
    s = o.SimulationObject(p, psi0)
    s.psik = o.forward(psi0)
    s.check_alias()
    alias = ctx.step(np.zeros(1), np.zeros(1))     # zero dt: psi_k unchanged, alias of the IC
    assert s.last_alias_mass > 0.1
    assert abs(alias[0] - s.last_alias_mass) < 1e-12 * s.last_alias_mass
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------
# host-logic level: trajectories -- rows a7, a8, a11, a13, a15, a16
# ---------------------------------------------------------------------------------------------------------------
def run_both(ps, steps, coupling=m.COUPLING_INDEPENDENT, chunk=0, psi0s=None):
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=len(ps), coupling=coupling, chunk_streams=chunk)
    refs = []
    for i, p in enumerate(ps):
        psi0 = initial_wavefunction(p) if psi0s is None else psi0s[i]
        sim.set_psi(i, psi0)
        refs.append(o.SimulationObject(p, psi0))
    worst = 0.0
    for k in range(steps):
        if not sim.not_finished():
            break
        sim.update()
        for i, r in enumerate(refs):
            if not r.not_finished():
                continue
            r.update()
            st = sim.state(i)
            # a potential-limited dt inherits the relative error of max|phi|; dump-limited ones are exact
            assert abs(st.dt - r.last_dt) <= 1e-12 * abs(r.last_dt), (k, i, st.dt, r.last_dt)
            assert abs(st.time - r.parameters.time) <= 1e-13 * abs(r.parameters.time)
            # on 4^3 / 8^3 grids the tophat's potential is a 1e-5 residual of cancelling terms (ill-conditioned)
            ptol = 1e-12 if ps[0].size >= 16 else 1e-9
            assert abs(st.potential_max - r.last_potential_max) <= ptol * r.last_potential_max
            assert alias_close(st.alias_mass, r.last_alias_mass)
            assert st.current_dumps == r.parameters.current_dumps and st.n_steps == r.parameters.n_steps
            if ps[0].expanding:
                assert abs(st.tau - r.parameters.tau) <= 1e-13 * abs(r.parameters.tau)
                assert abs(st.scale_factor - r.scale_factor_solver.get_a()) <= 1e-14
            if k % 4 == 3 or k == steps - 1:
                worst = max(worst, rel_l2(sim.get_psi(i), r.psi))
    return sim, refs, worst


@pytest.mark.parametrize("name,size,nstreams,steps", [
    ("spherical-tophat", None, 3, 12), ("spherical-tophat-cosmo", None, 3, 12), ("planeWave3d_e10_sym", None, 3, 12),
    ("spherical-tophat", 32, 2, 6), ("spherical-tophat", 64, 2, 4), ("gaussian-overdensity-mft", 64, 1, 4),
    ("repro-planeWave1d", None, 3, 8), ("spherical-tophat", 4, 2, 3), ("spherical-tophat", 8, 3, 3),
    ("spherical-tophat", 64, 3, 3), ("spherical-tophat", 128, 3, 2),      # several tiles per CTA, a one-stream group
    ("spherical-tophat-cosmo", 64, 3, 4), ("spherical-tophat-cosmo", 128, 2, 2),  # expanding box beyond 16^3
])
def test_trajectory_matches_oracle(name, size, nstreams, steps):
    ps = oracle_streams(name, size, limit=nstreams)
    sim, refs, worst = run_both(ps, steps)
    assert worst < 1e-10, worst
    sim.close()


def test_cuda_path_matches_the_reference_legacy_python_run():
    """tests/golden/legacy_grav3d_traj.npz: the reference's own legacy Python integrator (python_deprecated/gravSolver.py,
    `Grav3D.Update`) executed in the build container with a fixed dt -- 1, 10 and 40 steps of a collapsing Gaussian.  The
    CUDA path is compared with that run DIRECTLY (not through the oracle)."""
    z = np.load(os.path.join(GOLDEN, "legacy_grav3d_traj.npz"))
    n, length, hbar_, mtot, c, dt = z["params"]
    p = o.SimulationParameters(axis_length=float(length), time=0.0, final_sim_time=float(dt) * 40, cfl=1e9, num_data_dumps=40,
                               total_mass=float(mtot), particle_mass=o.HBAR / float(hbar_), sim_name="legacy",
                               k2_cutoff=0.95, alias_threshold=1e9, hbar_=float(hbar_), dims=3, size=int(n))
    sim = m.SimulationObject(to_msm_params(p), n_streams=2)
    sim.set_psi(0, z["psi0"])
    sim.set_psi(1, z["psi0"])
    assert rel_l2(sim.grid.get_potential(0), z["phi0"]) < 1e-12
    for k in range(1, 41):
        sim.update()
        assert abs(sim.state(0).dt - float(dt)) <= 1e-15
        if k in (1, 10, 40):
            assert rel_l2(sim.get_psi(0), z[f"psi_{k:03d}"]) < 1e-12, k
            assert rel_l2(sim.get_psi(1), z[f"psi_{k:03d}"]) < 1e-12, k
    sim.close()


def test_two_dimensional_grid():
    t = __import__("golden_util").load_toml("spherical-tophat", 32)
    t.dims = 2
    ps = list(o.simulation_iter(t))[:3]
    sim, refs, worst = run_both(ps, 5)
    assert worst < 1e-10
    sim.close()


@pytest.mark.parametrize("name", ["spherical-tophat", "spherical-tophat-cosmo", "planeWave3d_e10_sym"])
def test_against_committed_golden_trajectories(name):
    z = np.load(f"{GOLDEN}/traj_{name}.npz")
    its = oracle_streams(name)
    ps = [next(q for q in its if q.sim_name == str(s)) for s in z["streams"]]
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=len(ps))
    for j in range(len(ps)):
        sim.set_psi(j, z[f"s{j}_psi0"])
    for step in range(1, 26):
        sim.update()
        for j in range(len(ps)):
            row = z[f"s{j}_scalars"][step - 1]
            st = sim.state(j)
            assert abs(st.dt - row[0]) <= 1e-13 * abs(row[0])
            assert abs(st.potential_max - row[1]) <= 1e-12 * abs(row[1])
            assert alias_close(st.alias_mass, row[2])
            assert abs(st.time - row[3]) <= 1e-13 * abs(row[3])
            assert st.current_dumps == int(row[5])
            key = f"s{j}_psi_{step}"
            if key in z.files:
                assert rel_l2(sim.get_psi(j), z[key]) < 1e-10
    sim.close()


def test_full_run_to_final_time_with_dumps(tmp_path):
    """examples/spherical-tophat.toml to t = final (200 steps, 200 dumps) for one sampled stream + the MFT run:
    every dump compared; on-disk layout of utils/io.rs:34-88 checked."""
    its = oracle_streams("spherical-tophat")
    ps = [its[0], its[-1]]
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=2)
    refs = []
    for i, p in enumerate(ps):
        psi0 = initial_wavefunction(p)
        sim.set_psi(i, psi0)
        refs.append(o.SimulationObject(p, psi0))
    worst, steps = 0.0, 0
    while sim.not_finished():
        sim.update()
        steps += 1
        for i, r in enumerate(refs):
            r.update()
            st = sim.state(i)
            assert st.dumped == 1 and st.current_dumps == r.parameters.current_dumps
            if st.current_dumps % 20 == 0:
                worst = max(worst, rel_l2(sim.get_psi(i), r.psi))
                sim.dump(i, str(tmp_path), ps[i].sim_name, st.current_dumps)
    sim.wait_io()
    assert steps == 200 and not any(r.not_finished() for r in refs)
    assert sim.state(0).finished == 1 and sim.state(0).time == 40.0
    assert worst < 1e-10, worst
    d = tmp_path / ps[0].sim_name
    re = np.load(open(d / "psi_00200_real", "rb"))       # extension-less NPY, shape (n, n, n, 1), f64
    im = np.load(open(d / "psi_00200_imag", "rb"))
    assert re.shape == (16, 16, 16, 1) and re.dtype == np.float64
    assert rel_l2(re[..., 0] + 1j * im[..., 0], refs[0].psi) < 1e-10
    sim.close()


@pytest.mark.parametrize("nstreams,chunk", [(1, 0), (3, 2), (5, 4), (7, 2), (11, 8), (9, 0)])
def test_odd_stream_counts_and_chunking(nstreams, chunk):
    its = oracle_streams("spherical-tophat")
    ps = its[:nstreams]
    sim, refs, worst = run_both(ps, 3, chunk=chunk)
    assert worst < 1e-10
    sim.close()


def test_streams_finish_at_different_steps():
    """Streams with different adaptive dt take different numbers of steps per dump; finished streams are masked."""
    its = oracle_streams("gaussian-overdensity-mft", 16, limit=1)
    p = its[0]
    p.final_sim_time, p.num_data_dumps = 6.0, 3
    ps = [p, o.SimulationParameters(**{**{f: getattr(p, f) for f in ("axis_length", "time", "final_sim_time", "cfl",
          "num_data_dumps", "total_mass", "particle_mass", "k2_cutoff", "alias_threshold", "hbar_", "dims", "size")},
          "sim_name": "b", "ics": p.ics})]
    base = initial_wavefunction(p)
    psi_b = o.cold_gauss([15.0] * 3, [2.5] * 3, p)      # concentrated => deeper potential => smaller dt
    sim, refs, worst = run_both(ps, 200, psi0s=[base, psi_b])
    assert not sim.not_finished() and all(not r.not_finished() for r in refs)
    assert sim.state(0).n_steps == refs[0].parameters.n_steps != refs[1].parameters.n_steps == sim.state(1).n_steps
    for i, r in enumerate(refs):
        assert rel_l2(sim.get_psi(i), r.psi) < 1e-10
    sim.close()


# ---------------------------------------------------------------------------------------------------------------
# pipelined outer loop (simulator/src/main.rs:43-85): uploads / downloads of neighbouring groups overlap the steps
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nstreams,chunk,max_updates,lb", [(11, 2, 3, 0), (16, 4, 0, 0), (5, 8, 1, 0), (9, 4, 2, 2),
                                                           (1, 0, 2, 0)])
def test_run_streams_matches_oracle_and_the_sequential_path(monkeypatch, nstreams, chunk, max_updates, lb):
    if lb:
        monkeypatch.setenv("MSM_B200_LB", str(lb))
    its = oracle_streams("spherical-tophat")
    ps = [copy.deepcopy(its[i % len(its)]) for i in range(nstreams)]
    for p in ps:
        p.final_sim_time, p.num_data_dumps = 1.0, 5          # 5 updates to the end (every step hits a dump)
    psi0s = [initial_wavefunction(p) for p in ps]
    flat = [np.ascontiguousarray(a).reshape(-1).view(np.float64).copy() for a in psi0s]
    re = [np.zeros(psi0s[0].shape) for _ in ps]
    im = [np.zeros(psi0s[0].shape) for _ in ps]
    order = list(range(nstreams))[::-1] if nstreams % 2 else list(range(nstreams))      # any order of slots
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=nstreams, chunk_streams=chunk)
    sim.run_streams(order, [flat[s] for s in order], [re[s] for s in order], [im[s] for s in order], max_updates)
    for s, p in enumerate(ps):
        r = o.SimulationObject(copy.deepcopy(p), psi0s[s])     # the oracle advances its parameters in place
        k = 0
        while r.not_finished() and (max_updates == 0 or k < max_updates):
            r.update()
            k += 1
        st = sim.state(s)
        assert st.n_steps == k and (max_updates == 0 or k == max_updates)
        assert st.finished == (0 if r.not_finished() else 1)
        assert abs(st.time - r.parameters.time) <= 1e-13 * abs(r.parameters.time)
        assert rel_l2(re[s] + 1j * im[s], r.psi) < 1e-10
        assert rel_l2(sim.get_psi(s), re[s] + 1j * im[s]) < 1e-14       # the download is the resident state
    # continue the same streams without new ICs: state carries over
    if max_updates:
        sim.run_streams(order, None, [re[s] for s in order], None, 0)
        for s, p in enumerate(ps):
            r = o.SimulationObject(copy.deepcopy(p), psi0s[s])
            while r.not_finished():
                r.update()
            assert sim.state(s).finished == 1 and sim.state(s).n_steps == r.parameters.n_steps
            assert rel_l2(re[s], r.psi.real) < 1e-10
    # a second run from fresh ICs restarts time and dumps (a new SimulationObject per stream, :404-449)
    sim.run_streams(order[:1], [flat[order[0]]], None, None, 1)
    st = sim.state(order[0])
    assert st.n_steps == 1 and st.current_dumps == 1
    sim.close()


def test_run_streams_seeded_builds_the_initial_conditions_on_the_device():
    """msm_sim_run_streams_seeded: un-sampled IC saved once on the device + the seeded sampler per stream (the reference's
    `new_from_params`, simulation_object.rs:404-435), then the same pipelined loop; against the oracle run of every stream
    of examples/spherical-tophat.toml (Husimi seeds + the trailing un-sampled mean-field run)."""
    ps = oracle_streams("spherical-tophat")
    ps = ps[:5] + [ps[-1]]
    assert ps[-1].sampling_parameters is None
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=len(ps), chunk_streams=2)
    ics = ps[0].ics
    sim.grid.ic_spherical_tophat(0, ps[0].axis_length, float(ics["radius"]), float(ics["delta"]), float(ics["slope"]))
    sim.grid.ic_store(0)
    seeds = [p.sampling_parameters["seed"] if p.sampling_parameters else None for p in ps]
    re = [np.empty((16,) * 3) for _ in ps]
    im = [np.empty((16,) * 3) for _ in ps]
    sim.run_streams_seeded(list(range(len(ps))), "Husimi", seeds, re, im, max_updates=4)
    for i, p in enumerate(ps):
        ref = o.SimulationObject(p, initial_wavefunction(p))
        for _ in range(4):
            ref.update()
        assert sim.state(i).n_steps == 4 and abs(sim.state(i).time - ref.parameters.time) <= 1e-13 * ref.parameters.time
        assert rel_l2(re[i] + 1j * im[i], ref.psi) < 1e-10, i
    sim.close()


def test_async_transfers_order_against_compute():
    """msm_upload_begin / msm_download_begin: uploads land before the compute stream touches the stream, downloads
    see everything enqueued before them."""
    ps = oracle_streams("spherical-tophat", 32, limit=4)
    ctx = make_ctx(ps[0], 4, chunk_streams=2)
    assert ctx.chunk_streams() == 2
    psi0s = [initial_wavefunction(p) for p in ps]
    flat = [np.ascontiguousarray(a).reshape(-1).view(np.float64).copy() for a in psi0s]
    for s in range(4):
        ctx.upload_begin(s, flat[s])
    re = [np.zeros(psi0s[0].shape) for _ in ps]
    im = [np.zeros(psi0s[0].shape) for _ in ps]
    hb = ps[0].hbar_
    alias = ctx.step([0.05 / 4 * hb] * 4, [0.05 / hb] * 4)
    assert np.all(alias >= 0)
    for s in range(4):                                   # more downloads than staging buffers
        ctx.download_begin(s, re[s], im[s])
    ctx.upload_begin(0, flat[1])                         # overwriting stream 0 must wait for its download
    ctx.transfers_wait()
    want = []
    for s in range(4):
        ctx2 = make_ctx(ps[0], 2, chunk_streams=2)       # same pairing as above: (0,1) and (2,3)
        base = (s // 2) * 2
        ctx2.set_psi(0, psi0s[base])
        ctx2.set_psi(1, psi0s[base + 1])
        ctx2.step([0.05 / 4 * hb] * 2, [0.05 / hb] * 2)
        want.append(ctx2.get_psi(s - base))
        ctx2.close()
    for s in range(4):
        assert np.array_equal(re[s] + 1j * im[s], want[s])
    assert np.array_equal(ctx.get_psi(0), psi0s[1])
    ctx.close()


def test_run_streams_error_behaviour():
    ps = oracle_streams("spherical-tophat", limit=2)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=2)
    with pytest.raises(m.MsmError) as e:
        sim.run_streams([0, 0])
    assert e.value.code == _lib.MSM_E_ARG
    with pytest.raises(m.MsmError) as e:
        sim.run_streams([0, 1])                          # no wavefunction uploaded yet
    assert e.value.code == _lib.MSM_E_STATE
    sim.close()
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=2, coupling=m.COUPLING_SUMMED)
    with pytest.raises(m.MsmError) as e:
        sim.run_streams([0, 1])
    assert e.value.code == _lib.MSM_E_ARG
    sim.close()


# ---------------------------------------------------------------------------------------------------------------
# coupled (summed-density) mode -- north-star variant
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,nstreams", [("spherical-tophat", 4), ("spherical-tophat", 5), ("spherical-tophat-cosmo", 3)])
def test_summed_mode_matches_oracle(name, nstreams):
    ps = oracle_streams(name, limit=nstreams)
    psi0s = [initial_wavefunction(p) for p in ps]
    ens = o.SummedEnsemble(ps[0], psi0s)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=nstreams, coupling=m.COUPLING_SUMMED, chunk_streams=2)
    for i, a in enumerate(psi0s):
        sim.set_psi(i, a)
    for k in range(5):
        sim.update()
        ens.update()
        st = sim.state(0)
        assert abs(st.dt - ens.head.last_dt) <= 1e-13 * ens.head.last_dt
        assert abs(st.time - ens.head.parameters.time) <= 1e-13 * ens.head.parameters.time
    for i in range(nstreams):
        assert rel_l2(sim.get_psi(i), ens.streams[i].psi) < 1e-10
        assert alias_close(sim.state(i).alias_mass, ens.streams[i].last_alias_mass)
    sim.close()


@pytest.mark.parametrize("size,nstreams,chunk,lb,real,fuse", [(32, 3, 2, 0, 1, 1), (64, 2, 0, 0, 1, 1), (64, 5, 4, 2, 1, 1),
                                                              (128, 2, 0, 3, 1, 1), (32, 3, 2, 0, 0, 1), (32, 3, 2, 0, 1, 0)])
def test_summed_mode_real_field_solve(monkeypatch, size, nstreams, chunk, lb, real, fuse):
    """The shared potential of the summed-density mode is solved on the REAL density: n/2-point R2C / C2R passes along x,
    half-spectrum y / z passes, Nyquist plane (n/2 = 16 .. 64 here, 256 in test_gpu_headline.py), with the blocked
    device layout forced as well; real = 0 runs the complex fallback for comparison.  Trajectory, dt, max|phi| and the
    potential itself against the oracle ensemble."""
    monkeypatch.setenv("MSM_B200_LB", str(lb))
    monkeypatch.setenv("MSM_B200_REAL", str(real))
    monkeypatch.setenv("MSM_B200_FUSE", str(fuse))               # 0: un-fused pass sequences, no eager dt-potential
    ps = oracle_streams("spherical-tophat", size, limit=nstreams)
    psi0s = [initial_wavefunction(p) for p in ps]
    ens = o.SummedEnsemble(ps[0], psi0s)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=nstreams, coupling=m.COUPLING_SUMMED, chunk_streams=chunk)
    for i, a in enumerate(psi0s):
        sim.set_psi(i, a)
    for k in range(3):
        sim.update()
        ens.update()
        st = sim.state(nstreams - 1)
        assert abs(st.dt - ens.head.last_dt) <= 1e-13 * ens.head.last_dt
        assert abs(st.potential_max - ens.head.last_potential_max) <= 1e-12 * ens.head.last_potential_max
    for i in range(nstreams):
        assert rel_l2(sim.get_psi(i), ens.streams[i].psi) < 1e-10
        assert alias_close(sim.state(i).alias_mass, ens.streams[i].last_alias_mass)
    ens._potential()
    assert rel_l2(sim.grid.get_potential(0), ens.head.phi.real) < 1e-12
    sim.update()                                     # the cached max|phi| survived the potential download
    ens.update()
    assert abs(sim.state(0).dt - ens.head.last_dt) <= 1e-13 * ens.head.last_dt
    sim.close()


@pytest.mark.parametrize("dims", [1, 2])
def test_summed_mode_in_one_and_two_dimensions(dims):
    """dims < 3 keep the complex Poisson solve (the real-field path needs the three-pass structure)"""
    t = __import__("golden_util").load_toml("spherical-tophat", 32)
    t.dims = dims
    ps = list(o.simulation_iter(t))[:3]
    psi0s = [initial_wavefunction(p) for p in ps]
    ens = o.SummedEnsemble(ps[0], psi0s)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=3, coupling=m.COUPLING_SUMMED, chunk_streams=2)
    for i, a in enumerate(psi0s):
        sim.set_psi(i, a)
    for _ in range(4):
        sim.update()
        ens.update()
        assert abs(sim.state(1).dt - ens.head.last_dt) <= 1e-13 * ens.head.last_dt
    for i in range(3):
        assert rel_l2(sim.get_psi(i), ens.streams[i].psi) < 1e-10
    sim.close()


def test_summed_mode_with_identical_streams_equals_independent():
    p = oracle_streams("spherical-tophat")[-1]
    psi0 = initial_wavefunction(p)
    a = m.SimulationObject(to_msm_params(p), n_streams=4, coupling=m.COUPLING_SUMMED)
    b = m.SimulationObject(to_msm_params(p), n_streams=1)
    for i in range(4):
        a.set_psi(i, psi0)
    b.set_psi(0, psi0)
    for _ in range(4):
        a.update()
        b.update()
    assert rel_l2(a.get_psi(3), b.get_psi(0)) < 1e-12
    assert abs(a.state(2).dt - b.state(0).dt) <= 1e-13 * b.state(0).dt
    a.close()
    b.close()


# ---------------------------------------------------------------------------------------------------------------
# on-device initial conditions (row f-1) and errors
# ---------------------------------------------------------------------------------------------------------------
def test_device_initial_conditions_match_oracle():
    p = oracle_streams("spherical-tophat", limit=1)[0]
    ctx = make_ctx(p, 2)
    ctx.ic_spherical_tophat(0, p.axis_length, 5.0, 100.0, 50.0)
    base = o.spherical_tophat(p, 5.0, 100.0, 50.0)
    assert rel_l2(ctx.get_psi(0), base) < 1e-13
    ctx.ic_copy(1, 0)
    ctx.sample_perturbation(1, "Husimi", 7, p.n_tot)
    want = o.sample_quantum_perturbation(base, p, {"seed": 7, "scheme": "Husimi"})
    assert rel_l2(ctx.get_psi(1), want) < 1e-13
    ctx.sample_perturbation(0, "Wigner", 2 ** 40 + 3, p.n_tot)
    want = o.sample_quantum_perturbation(base, p, {"seed": 2 ** 40 + 3, "scheme": "Wigner"})
    assert rel_l2(ctx.get_psi(0), want) < 1e-13
    ctx.close()
    g = oracle_streams("gaussian-overdensity-mft", 32, limit=1)[0]
    ctx = make_ctx(g, 1)
    ctx.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    assert rel_l2(ctx.get_psi(0), o.cold_gauss([15.0] * 3, [10.0] * 3, g)) < 1e-13
    # cold_gauss_kspace (ics.rs:282-431): Gaussian in k, random phases, forward transform
    ctx.ic_cold_gauss_kspace(0, [0.1, 0.0, -0.1], [0.3, 0.25, 0.2], phase_seed=5)
    want = o.cold_gauss_kspace([0.1, 0.0, -0.1], [0.3, 0.25, 0.2], g, 5)
    assert rel_l2(ctx.get_psi(0), want) < 1e-13 and o.check_norm(want, g.dx, 3)
    ctx.close()


def test_device_poisson_sampling_scheme_statistics():
    """ics.rs:495-558: |psi| <- sqrt(Pois(|psi|^2 dV n_tot) / n_tot), phase kept.  The reference's draw is unseeded
    (thread_rng), so the check is distributional: mean and variance of the counts equal lambda (both branches of the
    generator: product method below 10, PTRS above), phases survive, seeds decorrelate, empty cells stay empty."""
    p = oracle_streams("spherical-tophat", 32, limit=1)[0]
    dv, cells = p.dx ** 3, 32 ** 3
    rng = np.random.default_rng(3)
    phase = np.exp(2j * np.pi * rng.random((32, 32, 32)))
    for lam in (0.4, 6.0, 35.0, 4.0e4, 3.0e9):
        n_tot = 1.0e12
        amp = np.sqrt(lam / (dv * n_tot))
        psi0 = amp * phase
        psi0[0, 0, :4] = 0.0                                       # empty cells
        ctx = make_ctx(p, 2)
        ctx.set_psi(0, psi0)
        ctx.set_psi(1, psi0)
        ctx.sample_perturbation(0, "Poisson", 11, n_tot)
        ctx.sample_perturbation(1, "Poisson", 12, n_tot)
        a, b = ctx.get_psi(0), ctx.get_psi(1)
        ctx.close()
        counts = np.abs(a) ** 2 * dv * n_tot
        assert np.all(counts[0, 0, :4] == 0.0)
        live = np.ones(counts.shape, bool)
        live[0, 0, :4] = False
        c = counts[live]
        assert np.allclose(c, np.rint(c), rtol=0, atol=1e-6 * max(1.0, lam))             # integer counts
        assert abs(c.mean() - lam) < 6.0 * np.sqrt(lam / c.size)
        assert abs(c.var() - lam) < 6.0 * lam * np.sqrt(2.0 / c.size) + 6.0 * np.sqrt(lam / c.size)
        nz = live & (counts > 0)
        assert np.max(np.abs(a[nz] / np.abs(a[nz]) - phase[nz])) < 1e-12                 # phases kept
        cb = (np.abs(b) ** 2 * dv * n_tot)[live]
        assert abs(np.corrcoef(c, cb)[0, 1]) < 6.0 / np.sqrt(c.size)                     # another seed, another draw
    # same seed -> same field, on any chunking
    ctx = make_ctx(p, 1)
    ctx.set_psi(0, psi0)
    ctx.sample_perturbation(0, "Poisson", 11, n_tot)
    assert np.array_equal(ctx.get_psi(0), a)
    ctx.close()


def test_error_behaviour():
    p = oracle_streams("spherical-tophat", limit=1)[0]
    ctx = make_ctx(p, 2)
    with pytest.raises(m.MsmError) as ei:                  # no psi yet
        ctx.potential_max()
    assert ei.value.code == _lib.MSM_E_STATE
    with pytest.raises(m.MsmError) as ei:
        ctx.get_psi(5)
    assert ei.value.code == _lib.MSM_E_ARG
    ctx.close()
    # aliasing is reported per stream, the reference panics (simulation_object.rs:607-617)
    p.alias_threshold = 1e-40
    rng = np.random.default_rng(1)
    noisy = o.normalize(rng.standard_normal((16,) * 3) + 1j * rng.standard_normal((16,) * 3), p.dx, 3)
    sim = m.SimulationObject(to_msm_params(p), n_streams=1)
    sim.set_psi(0, noisy)
    with pytest.raises(m.FourierAliasing):
        sim.update()
    assert sim.state(0).aliased == 1
    ref = o.SimulationObject(p, noisy)
    with pytest.raises(o.FourierAliasing) as oe:
        ref.update()
    assert abs(sim.state(0).alias_mass - oe.value.p_mass) <= 1e-10 * oe.value.p_mass
    sim.close()


def test_driver_runs_a_reference_toml(tmp_path):
    """python -m msm_b200 --toml <file>: the host driver end to end, dumps in the reference layout."""
    import shutil
    from msm_b200 import driver
    from msm_b200.config import read_toml
    toml = tmp_path / "run.toml"
    toml.write_text("""
axis_length = 30
final_sim_time = 1.0
cfl = 0.5
num_data_dumps = 5
total_mass = 1e11
hbar_ = 0.05
sim_name = "spherical-tophat"
k2_cutoff = 0.95
alias_threshold = 0.02
dims = 3
size = 16
[ics]
type = "SphericalTophat"
radius = 5.0
slope = 50
delta = 100
[sampling]
seeds = "1 to 3"
scheme = "Husimi"
""")
    cfg = read_toml(str(toml))
    res = driver.run(cfg, out_root=str(tmp_path / "sim-data"))
    assert res["streams"] == 4 and res["stream_steps"] == 20
    ot = o.read_toml(str(toml))
    for p in o.simulation_iter(ot):
        ref = o.run_stream(p, o.initial_wavefunction(p))
        for idx, psi in ref.dumps:
            d = tmp_path / "sim-data" / p.sim_name
            got = np.load(open(d / f"psi_{idx:05d}_real", "rb"))[..., 0] + 1j * np.load(open(d / f"psi_{idx:05d}_imag", "rb"))[..., 0]
            assert rel_l2(got, psi) < 1e-10, (p.sim_name, idx)


# ---------------------------------------------------------------------------------------------------------------
# device layout: the blocked slow axis (used for n >= 512) forced on small grids, where the oracle is cheap
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lb,size", [(2, 16), (3, 32), (1, 64), (4, 16)])
def test_blocked_device_layout_on_small_grids(monkeypatch, lb, size):
    monkeypatch.setenv("MSM_B200_LB", str(lb))
    rng = np.random.default_rng(size + lb)
    a = rng.standard_normal((3, size, size, size)) + 1j * rng.standard_normal((3, size, size, size))
    assert rel_l2(m.forward(a, 3), sf.fftn(a, axes=(1, 2, 3), norm="ortho")) < 1e-13
    ps = oracle_streams("spherical-tophat", size, limit=3)
    sim, refs, worst = run_both(ps, 4)
    assert worst < 1e-10
    # on-device ICs and the sampler address cells through the same mapping
    p = ps[0]
    ctx = make_ctx(p, 1)
    ctx.ic_spherical_tophat(0, p.axis_length, 5.0, 100.0, 50.0)
    ctx.sample_perturbation(0, "Wigner", 5, p.n_tot)
    want = o.sample_quantum_perturbation(o.spherical_tophat(p, 5.0, 100.0, 50.0), p, {"seed": 5, "scheme": "Wigner"})
    assert rel_l2(ctx.get_psi(0), want) < 1e-13
    assert rel_l2(ctx.get_potential(0), _oracle_potential(p, want)) < 1e-12
    re, im = ctx.get_psi_planes(0)
    assert rel_l2(re + 1j * im, want) < 1e-13
    assert rel_l2(ctx.get_psik(0), o.forward(want)) < 1e-13
    ctx.close()
    sim.close()


def _oracle_potential(p, psi):
    s = o.SimulationObject(p, psi)
    s.calculate_potential()
    return s.phi.real


# ---------------------------------------------------------------------------------------------------------------
# row f-3: the synthesizer's stream reductions on the device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lb", [0, 2])
def test_ensemble_statistics_match_synthesizer_restatement(monkeypatch, tmp_path, lb):
    from msm_b200 import driver
    monkeypatch.setenv("MSM_B200_LB", str(lb))
    ps = oracle_streams("spherical-tophat", limit=5)
    sim, refs, worst = run_both(ps, 3, chunk=2)
    want = o.synthesizer_combine([r.psi for r in refs], ps[0].dx ** 3)
    got = driver.combine_streams(sim, 5, ps[0].dx)
    for k in ("psi", "psi2", "psik", "psik2"):
        assert rel_l2(got[k], want[k]) < 1e-12, k
        if k in ("psi2", "psik2"):
            assert np.max(np.abs(got[k].imag)) == 0.0
    # Qx is a difference of O(1) terms (it is ~0 for these nearly identical streams): compare on that scale
    assert abs(got["Qx"] - want["Qx"]) <= 1e-12 * abs(np.sum(want["psi2"]) * ps[0].dx ** 3)
    # a subset of streams, and the on-disk layout of dump_complex
    sub = driver.combine_streams(sim, 2, ps[0].dx, active=[1, 0, 0, 1, 0])
    want2 = o.synthesizer_combine([refs[0].psi, refs[3].psi], ps[0].dx ** 3)
    assert rel_l2(sub["psik2"], want2["psik2"]) < 1e-12
    driver.write_combined(str(tmp_path), "spherical-tophat", 3, got, 3, 16)
    re = np.load(open(tmp_path / "spherical-tophat-combined" / "psik_00003_real", "rb"))
    assert re.shape == (16, 16, 16, 1) and rel_l2(re[..., 0], want["psik"].real) < 1e-12
    q = np.load(open(tmp_path / "spherical-tophat-combined" / "Qx_00003_real", "rb"))
    assert q.shape == (1, 1, 1, 1)
    sim.close()


# ---------------------------------------------------------------------------------------------------------------
# the native (C++) host over the C ABI: msm_b200/msm-simulator-b200, mirror of simulator/src/main.rs
# ---------------------------------------------------------------------------------------------------------------
def test_native_host_runs_a_reference_config(tmp_path):
    import subprocess
    from msm_b200 import driver
    from msm_b200.config import read_toml
    from conftest import ROOT
    toml = tmp_path / "run.toml"
    toml.write_text("""
axis_length = 30
final_sim_time = 0.8
cfl = 0.5
num_data_dumps = 4
total_mass = 1e11
hbar_ = 0.05
sim_name = "tophat-native"
k2_cutoff = 0.95
alias_threshold = 0.02
dims = 3
size = 16
[ics]
type = "SphericalTophat"
radius = 5.0
slope = 50
delta = 100
[sampling]
seeds = "3 to 4"
scheme = "Wigner"
""")
    cfg = read_toml(str(toml))
    params = tmp_path / "run.params"
    driver.export_params(cfg, str(params))
    exe = os.path.join(ROOT, "msm_b200", "msm-simulator-b200")
    out = subprocess.run([exe, "--params", str(params), "--out", str(tmp_path / "sim-data"), "--verbose"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "3 streams" in out.stdout
    for p in o.simulation_iter(o.read_toml(str(toml))):
        ref = o.run_stream(p, o.initial_wavefunction(p))
        assert len(ref.dumps) == 5
        for idx, psi in ref.dumps:
            d = tmp_path / "sim-data" / p.sim_name
            got = np.load(open(d / f"psi_{idx:05d}_real", "rb"))[..., 0] + 1j * np.load(open(d / f"psi_{idx:05d}_imag", "rb"))[..., 0]
            assert rel_l2(got, psi) < 1e-10, (p.sim_name, idx)
