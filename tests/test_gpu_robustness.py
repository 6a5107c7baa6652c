"""Error paths and the dump pipeline of the CUDA path (through the C ABI): NaN / Inf guard (utils/grid.rs:66-105,
RuntimeError::NanOrInf), dump writer failures (RuntimeError::IOError), potential dumps (simulation_object.rs:1167-1180),
the cached max|phi| after a non-blocking step, and dumps that do not stall the step loop."""
import os
import time

import numpy as np
import pytest

import msm_b200 as m
from msm_b200 import _lib
from oracle import msm_oracle as o
from conftest import rel_l2
from golden_util import initial_wavefunction, oracle_streams, to_msm_params

pytestmark = pytest.mark.gpu


def oracle_potential(p, psi):
    ref = o.SimulationObject(p, psi)
    ref.calculate_potential()
    return ref.phi.real


def test_potential_max_after_a_nonblocking_step_and_a_new_wavefunction():
    """msm_step(alias_mass = NULL) leaves max|phi| of the stepped psi waiting on the device; a wavefunction uploaded
    afterwards must not inherit it (the adaptive dt would silently be wrong)."""
    ps = oracle_streams("spherical-tophat", 32, limit=3)
    osim = o.SimulationObject(ps[0], np.zeros((2, 2, 2), dtype=np.complex128))
    ctx = m.Context(3, 32, 2, ps[0].dx, osim.density_prefactor(), osim.poisson_coeff(), ps[0].k2_cutoff)
    psi = [initial_wavefunction(p) for p in ps]
    ctx.set_psi(0, psi[0])
    ctx.set_psi(1, psi[1])
    ctx.potential_max()
    assert ctx.step(np.full(2, 1e-3), np.full(2, 1e-2), blocking=False) is None
    rng = np.random.default_rng(5)
    new = o.normalize(psi[2] * np.exp(1j * rng.standard_normal(psi[2].shape)) * (1.0 + rng.random(psi[2].shape)), ps[0].dx, 3)
    ctx.set_psi(1, new)                                       # stream 0 keeps its pending value, stream 1 must not
    got = ctx.potential_max()
    want = np.abs(oracle_potential(ps[0], new)).max()
    assert abs(got[1] - want) <= 1e-12 * want, (got[1], want)
    ctx.read_alias()
    ctx.close()


@pytest.mark.parametrize("bad", [np.nan, np.inf])
def test_nan_or_inf_is_reported_per_stream(bad):
    ps = oracle_streams("spherical-tophat", limit=2)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=2)
    good = initial_wavefunction(ps[0])
    broken = good.copy()
    broken[3, 4, 5] = complex(bad, 0.0)
    sim.set_psi(0, good)
    sim.set_psi(1, broken)
    with pytest.raises(m.MsmError) as e:
        sim.update()
    assert e.value.code == _lib.MSM_E_NAN and "stream 1" in e.value.msg
    # the healthy stream alone still runs
    sim.update_streams([1, 0])
    assert sim.state(0).n_steps == 1 and sim.state(1).n_steps == 0
    sim.close()


def test_nan_inside_the_step_is_reported():
    """an Inf kick coefficient turns psi into NaN inside msm_step: the alias sum carries it to the host"""
    p = oracle_streams("spherical-tophat", limit=1)[0]
    osim = o.SimulationObject(p, np.zeros((2, 2, 2), dtype=np.complex128))
    ctx = m.Context(3, 16, 1, p.dx, osim.density_prefactor(), osim.poisson_coeff(), p.k2_cutoff)
    ctx.set_psi(0, initial_wavefunction(p))
    with pytest.raises(m.MsmError) as e:
        ctx.step([1e-3], [np.inf])
    assert e.value.code == _lib.MSM_E_NAN
    ctx.close()


def test_potential_dump_matches_the_oracle(tmp_path):
    """output_potential (simulation_object.rs:1167-1180): potential_%05d_real holds phi, potential_%05d_imag zeros."""
    ps = oracle_streams("spherical-tophat", limit=2)
    sim = m.SimulationObject(to_msm_params(ps[0]), n_streams=2)
    refs = []
    for i, p in enumerate(ps):
        a = initial_wavefunction(p)
        sim.set_psi(i, a)
        refs.append(o.SimulationObject(p, a))
    for _ in range(2):
        sim.update()
        for r in refs:
            r.update()
    for i, p in enumerate(ps):
        sim.dump(i, str(tmp_path), p.sim_name, 2)
        sim.dump_potential(i, str(tmp_path), p.sim_name, 2)
    sim.wait_io()
    for i, p in enumerate(ps):
        d = tmp_path / p.sim_name
        re = np.load(open(d / "potential_00002_real", "rb"))
        im = np.load(open(d / "potential_00002_imag", "rb"))
        assert re.shape == (16, 16, 16, 1) and re.dtype == np.float64 and not im.any()
        refs[i].calculate_potential()                         # :1168
        assert rel_l2(re[..., 0], refs[i].phi.real) < 1e-12
        psi = np.load(open(d / "psi_00002_real", "rb"))[..., 0] + 1j * np.load(open(d / "psi_00002_imag", "rb"))[..., 0]
        assert rel_l2(psi, refs[i].psi) < 1e-10
    sim.close()


def test_driver_writes_potential_dumps(tmp_path):
    from msm_b200 import driver
    from msm_b200.config import read_toml
    toml = tmp_path / "run.toml"
    toml.write_text("""
axis_length = 30
final_sim_time = 0.4
cfl = 0.5
num_data_dumps = 2
total_mass = 1e11
hbar_ = 0.05
sim_name = "pot"
k2_cutoff = 0.95
alias_threshold = 0.02
dims = 3
size = 16
output_potential = true
[ics]
type = "SphericalTophat"
radius = 5.0
slope = 50
delta = 100
""")
    cfg = read_toml(str(toml))
    assert cfg.output_potential
    driver.run(cfg, out_root=str(tmp_path / "sim-data"))
    p = list(o.simulation_iter(o.read_toml(str(toml))))[0]
    ref = o.run_stream(p, o.initial_wavefunction(p))
    ref.calculate_potential()
    d = tmp_path / "sim-data" / "pot"
    assert sorted(os.listdir(d)) == sorted(f"{f}_{i:05d}_{part}" for f in ("psi", "potential") for i in range(3)
                                            for part in ("real", "imag"))
    assert rel_l2(np.load(open(d / "potential_00002_real", "rb"))[..., 0], ref.phi.real) < 1e-12


def test_dump_writer_failure_is_an_io_error(tmp_path):
    """RuntimeError::IOError (utils/error.rs:5-27): the reference panics in the writer; here the failure surfaces as
    MSM_E_IO from msm_sim_wait_io / the next dump instead of being dropped."""
    p = oracle_streams("spherical-tophat", limit=1)[0]
    sim = m.SimulationObject(to_msm_params(p), n_streams=1)
    sim.set_psi(0, initial_wavefunction(p))
    blocker = tmp_path / "not-a-directory"
    blocker.write_text("x")
    with pytest.raises(m.MsmError) as e:                       # the directory cannot be created
        sim.dump(0, str(blocker), "run", 0)
    assert e.value.code == _lib.MSM_E_IO
    d = tmp_path / "out" / "run"
    d.mkdir(parents=True)
    (d / "psi_00001_real").mkdir()                             # the writer thread cannot open its file
    sim.dump(0, str(tmp_path / "out"), "run", 1)
    with pytest.raises(m.MsmError) as e:
        sim.wait_io()
    assert e.value.code == _lib.MSM_E_IO and "psi_00001_real" in e.value.msg
    sim.dump(0, str(tmp_path / "out"), "run", 2)               # reported once; later dumps work
    sim.wait_io()
    assert (d / "psi_00002_imag").exists() and (d / "psi_00001_imag").exists()
    sim.close()


def test_dumps_do_not_stall_the_step_loop(tmp_path, monkeypatch):
    """A dump after EVERY update (one stream of eight per update, 256^3: 256 MiB each) against the same loop without
    dumps.  msm_sim_dump only enqueues the inverse transform / plane split (compute stream) and the D2H copy (copy
    stream, pinned staging pool); NPY files are written by background threads.  The loop may cost the extra transform
    (4 of ~120 passes per update, plus its share of HBM and host bandwidth) but must not wait for PCIe or the disk
    (measured: +10 %; the bound is 25 % + 5 ms).  The pinned staging buffers are reserved up front."""
    monkeypatch.setenv("MSM_B200_DUMP_BUFFERS", "12")
    size, S, K = 256, 8, 12
    p = oracle_streams("gaussian-overdensity-mft", size, limit=1)[0]
    root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else str(tmp_path)
    root = os.path.join(root, f"msm_b200_dump_test_{os.getpid()}")
    sim = m.SimulationObject(to_msm_params(p), n_streams=S)
    g = sim.grid
    g.ic_cold_gauss(0, [15.0] * 3, [10.0] * 3)
    for s in range(1, S):
        g.ic_copy(s, 0)
    for s in range(S):
        g.sample_perturbation(s, "Wigner", s + 1, 1e10)
    try:
        def loop(dump, first_index):
            g.synchronize()
            t0 = time.perf_counter()
            for k in range(K):
                sim.update()
                if dump:
                    sim.dump(k % S, root, "run", first_index + k)
            g.synchronize()                                     # compute stream only: writers may still be busy
            return time.perf_counter() - t0
        sim.reserve_dump_buffers(12)                            # pinning host memory costs ~0.15 s per 256 MiB: up front
        loop(True, 0)                                           # warm-up: device staging, copy stream, first files
        sim.wait_io()
        plain = min(loop(False, 0) for _ in range(2))
        dumped = loop(True, 100)
        sim.wait_io()
        files = os.listdir(os.path.join(root, "run"))
        assert len(files) == 2 * 2 * K
        re = np.load(open(os.path.join(root, "run", f"psi_{100 + K - 1:05d}_real"), "rb"))
        assert re.shape == (size, size, size, 1) and np.isfinite(re).all()
        print(f"loop with a dump per update {dumped * 1e3:.1f} ms, without {plain * 1e3:.1f} ms")
        # measured +10 % (142 vs 129 ms: the extra inverse transform + plane split); a stalled loop was +90 %, a pinned
        # allocation inside the loop +100 ms each.  The bound leaves room for
        # a noisy box: this is a regression guard for "waits for PCIe / the disk", not a benchmark.
        assert dumped <= 1.25 * plain + 5e-3, (dumped, plain)
    finally:
        sim.close()
        import shutil
        shutil.rmtree(root, ignore_errors=True)
