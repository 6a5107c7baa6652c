"""Pins the oracle against every golden value the reference's own tests hold for this path,
plus closed-form known answers for the operators the reference never tests (drift, Poisson)."""
import math
import os

import numpy as np
import pytest

from oracle import msm_oracle as o
from conftest import rel_l2
from golden_util import GOLDEN, load_toml


def test_k_grid_golden():
    # simulator/src/utils/fft.rs:164-167 (and :176-179 for f64)
    assert np.array_equal(o.get_kgrid(0.25, 4), np.array([0.0, 1.0, -2.0, -1.0]))


def test_k_grid_requires_even():
    with pytest.raises(AssertionError):     # fft.rs:105
        o.get_kgrid(0.25, 5)


def test_spec_grid_golden():
    # simulator/src/utils/fft.rs:185-214 / :219-246: element q = i + j*size + k*size^2 of the linear buffer
    size, dims = 4, 3
    kg = o.get_kgrid(0.25, 4)
    values = np.zeros(size ** dims)
    for i in range(size):
        for j in range(size):
            for k in range(size):
                q = i + j * size + k * size * size
                values[q] = (kg[i] * kg[i] + kg[j] * kg[j] + kg[k] * kg[k]) * (2.0 * math.pi) ** 2.0
    host = o.spec_grid(0.25, dims, size).reshape(-1)
    assert np.array_equal(values, host)          # the reference asserts exact equality


@pytest.mark.parametrize("K,S,value,length", [
    (1, 8, complex(0.0, 128.0 ** -0.5), 128.0),   # tests/fft.rs:2-64 (norm checked with dk = dx)
    (1, 2, complex(0.0, 2.0), None),              # tests/fft.rs:66-... the (0, 2i) constant arrays
    (2, 2, complex(0.0, 2.0), None),
    (3, 2, complex(0.0, 2.0), None),
    (2, 8, complex(0.0, 2.0), None),
    (3, 8, complex(0.0, 2.0), None),
])
def test_fft_round_trip_and_unitarity(K, S, value, length):
    values = np.full((S,) * K, value, dtype=np.complex128)
    fk = o.forward(values)
    if length is not None:
        dx = length / S
        assert o.check_norm(values, dx, K)
        assert o.check_norm(fk, dx, K)            # dk = dx: unitary normalisation (tests/fft.rs:44-50)
    assert abs(np.sum(np.abs(fk) ** 2) - np.sum(np.abs(values) ** 2)) < 1e-12 * values.size
    back = o.inverse(fk)
    assert np.sum(np.abs(back - values)) < 1e-6   # reference's epsilon
    assert np.sum(np.abs(back - values)) < 1e-12


@pytest.mark.parametrize("D", [1, 2, 3])
def test_normalize(D):
    # utils/grid.rs:107-186: S = 8, dx = 1/S, values (1 + i)
    S = 8
    dx = 1.0 / S
    a = np.full((S,) * D, 1.0 + 1.0j)
    a = o.normalize(a, dx, D)
    assert abs(np.sum(np.abs(a) ** 2) * dx ** D - 1.0) < 1e-6
    assert o.check_norm(a, dx, D)


def test_parse_seeds():
    # common/src/parameters.rs:121-144
    assert o.parse_seeds("0..=55") == list(range(0, 56))
    assert o.parse_seeds("0 to 55") == list(range(0, 56))
    assert o.parse_seeds("[1, 3]") == [1, 3]
    assert o.parse_seeds("1, 3") == [1, 3]


def test_deserialize_toml(tmp_path):
    # simulator/src/utils/io.rs:248-326
    text = """
axis_length = 30.0
final_sim_time = 400.0
cfl = 0.5
num_data_dumps = 100
total_mass = 1e10
hbar_ = 0.02
sim_name = "gaussian-overdensity-512-mft"
k2_cutoff = 0.95
alias_threshold = 0.02
dims = 3
size = 512

[ics]
type = "ColdGaussKSpace"
mean = [15.0, 15.0, 15.0]
std = [10.0, 10.0, 10.0]

[sampling]
seeds = "1..=64"
scheme = "Husimi"
"""
    f = tmp_path / "t.toml"
    f.write_text(text)
    t = o.read_toml(str(f))
    assert (t.axis_length, t.final_sim_time, t.cfl, t.num_data_dumps, t.total_mass) == (30.0, 400.0, 0.5, 100, 1e10)
    assert t.hbar_ == 0.02 and t.sim_name == "gaussian-overdensity-512-mft"
    assert (t.k2_cutoff, t.alias_threshold, t.dims, t.size) == (0.95, 0.02, 3, 512)
    assert t.ics == {"type": "ColdGaussKSpace", "mean": [15.0] * 3, "std": [10.0] * 3}
    assert t.sampling == {"scheme": "Husimi", "seeds": list(range(1, 65))}


def test_ic_fixture_not_normalised():
    # initial_conditions/planeWave3d_e10_sym.npz: f64 (16,16,16), sum |psi|^2 = n/L = 0.26667
    z = np.load(GOLDEN + "/planeWave3d_e10_sym_ic.npz")
    assert z["real"].shape == (16, 16, 16) and z["real"].dtype == np.float64
    assert abs(np.sum(z["real"] ** 2 + z["imag"] ** 2) - 16.0 / 60.0) < 1e-4


def test_stream_iteration_order_and_names():
    # utils/io.rs:164-245: seeds ascending as "<sim>-stream%05d", then ONE un-sampled run "<sim>"
    its = list(o.simulation_iter(load_toml("spherical-tophat")))
    assert len(its) == 11
    assert [p.sim_name for p in its[:2]] == ["spherical-tophat-stream00001", "spherical-tophat-stream00002"]
    assert its[-1].sim_name == "spherical-tophat" and its[-1].sampling_parameters is None
    assert its[0].sampling_parameters == {"seed": 1, "scheme": "Husimi"}


def test_derived_parameters():
    # simulation_object.rs:243-274: hbar_, dx, dk = dx, k2_max = d pi^2 / dx^2
    p = list(o.simulation_iter(load_toml("spherical-tophat")))[0]
    assert p.dx == 30.0 / 16 and p.dk == p.dx
    assert abs(p.k2_max - 3 * math.pi ** 2 / p.dx ** 2) < 1e-12 * p.k2_max
    pc = list(o.simulation_iter(load_toml("spherical-tophat-cosmo")))[0]
    H0 = 0.7 * 1.022e-4
    Lc = math.sqrt(math.sqrt(1.5 * 0.7 * H0 ** 2) / 0.05) * 30.0 * 2.0       # Appendix A
    assert abs(pc.comoving_boxsize - Lc) < 1e-13 * Lc and pc.dx == pc.comoving_boxsize / 16


# ---- closed-form known answers (the reference has no tests for these operators) ---------------------------------
def _free_params(n=16, dims=3, mass=1e-30):
    return o.SimulationParameters(axis_length=10.0, time=0.0, final_sim_time=1.0, cfl=0.1, num_data_dumps=1,
                                  total_mass=mass, particle_mass=1.0, sim_name="kat", k2_cutoff=0.95,
                                  alias_threshold=1.0, hbar_=0.3, dims=dims, size=n)


def test_drift_plane_wave_known_answer():
    """With negligible mass the step is a free drift: psi = exp(i k.x) -> psi * exp(-i hbar_ k^2 t / 2)
    (two half drifts exp(-i dt hbar_ k^2 / 4), simulation_object.rs:504-516 and :562-574)."""
    p = _free_params()
    n, L = p.size, p.axis_length
    x = (np.arange(n) + 0.5) * p.dx
    m = (2, -1, 3)
    kvec = [2.0 * math.pi * mi / L for mi in m]
    psi0 = np.exp(1j * (kvec[0] * x[:, None, None] + kvec[1] * x[None, :, None] + kvec[2] * x[None, None, :]))
    psi0 = o.normalize(psi0, p.dx, 3)
    sim = o.SimulationObject(p, psi0)
    sim.update()
    t = sim.parameters.time
    k2 = sum(k * k for k in kvec)
    expect = psi0 * np.exp(-1j * p.hbar_ * k2 * t / 2.0)
    assert np.linalg.norm(sim.psi - expect) / np.linalg.norm(expect) < 1e-12


def test_poisson_single_mode_known_answer():
    """rho = A (1 + eps cos(k x)) / V  ->  phi = -C A eps cos(k x) / (V k^2), DC removed
    (simulation_object.rs:1066-1110 with c = -POIS_CONST)."""
    p = _free_params(n=32, mass=2.5e9)
    n, L = p.size, p.axis_length
    x = (np.arange(n) + 0.5) * p.dx
    k = 2.0 * math.pi * 3 / L
    eps = 0.2
    dens = (1.0 + eps * np.cos(k * x))[None, None, :] * np.ones((n, n, 1)) / L ** 3
    sim = o.SimulationObject(p, np.sqrt(dens).astype(np.complex128))
    sim.calculate_potential()
    expect = -o.POIS_CONST * p.total_mass * eps * np.cos(k * x)[None, None, :] / (L ** 3 * k * k) * np.ones((n, n, 1))
    assert np.max(np.abs(sim.phi.imag)) == 0.0
    assert np.linalg.norm(sim.phi.real - expect) / np.linalg.norm(expect) < 1e-12


def test_norm_is_conserved_by_the_step():
    ps = list(o.simulation_iter(load_toml("spherical-tophat")))
    from golden_util import initial_wavefunction
    sim = o.SimulationObject(ps[0], initial_wavefunction(ps[0]))
    for _ in range(5):
        sim.update()
    assert abs(np.sum(np.abs(sim.psi) ** 2) * ps[0].dx ** 3 - 1.0) < 1e-12


def test_rk4_and_scale_factor_known_answers():
    # utils/mod.rs:14-43 rk4 on y' = y ; Einstein-de Sitter a(t) = (a0^1.5 + 1.5 H0 t)^(2/3) for Om = 1
    y = 1.0
    for i in range(10):
        y = o.rk4(lambda t, yy: yy, i * 0.1, y, 0.1)
    assert abs(y - math.e) < 5e-6          # RK4 global error at h = 0.1
    c = o.CosmologyParameters(1.0, 0.0, 0.7, 9.0, 1e-3)
    s = o.ScaleFactorSolver(c)
    t = 500.0
    a = s.step(t)
    H0 = 0.7 * o.LITTLE_H_TO_BIG_H
    assert abs(a - (0.1 ** 1.5 + 1.5 * H0 * t) ** (2.0 / 3.0)) < 1e-10
    assert abs(s.get_time() - t) < 1e-9


def test_get_tau_eds_known_answer():
    # dtau/dt = sqrt(1.5 Om H0^2) / a^2 with a = (a0^1.5 + 1.5 H0 t)^(2/3): tau = sqrt(1.5) * 2 * (a0^-0.5 - a^-0.5)
    c = o.CosmologyParameters(1.0, 0.0, 0.7, 9.0, 0.01)
    t = 300.0
    H0 = 0.7 * o.LITTLE_H_TO_BIG_H
    a0 = 0.1
    a = (a0 ** 1.5 + 1.5 * H0 * t) ** (2.0 / 3.0)
    expect = math.sqrt(1.5) * 2.0 * (a0 ** -0.5 - a ** -0.5)
    assert abs(o.get_tau(t, c) - expect) < 1e-8 * expect


def test_synthesizer_combine_known_answers():
    """synthesizer/src/main.rs:63-93,161-173: identical streams have Qx = 0; psik is the UN-normalised DFT
    (psik[0] = sum psi); two streams psi and -psi average to zero with Qx = sum|psi|^2 dv."""
    rng = np.random.default_rng(0)
    psi = rng.standard_normal((8, 8, 8)) + 1j * rng.standard_normal((8, 8, 8))
    dv = 0.3 ** 3
    c = o.synthesizer_combine([psi, psi, psi], dv)
    assert abs(c["Qx"]) < 1e-12 and np.allclose(c["psi"], psi)
    assert abs(c["psik"][0, 0, 0] - psi.sum()) < 1e-10
    assert abs(np.sum(c["psik2"]).real - 8 ** 3 * np.sum(np.abs(psi) ** 2)) < 1e-7     # Parseval, un-normalised
    c = o.synthesizer_combine([psi, -psi], dv)
    assert np.max(np.abs(c["psi"])) == 0.0
    assert abs(c["Qx"] - np.sum(np.abs(psi) ** 2) * dv) < 1e-10


# ---------------------------------------------------------------------------------------------------------------
# the integrator step against the reference's own legacy Python integrator (python_deprecated/gravSolver.py)
# ---------------------------------------------------------------------------------------------------------------
def legacy_parameters():
    """tests/golden/legacy_grav3d_traj.npz: `Grav3D.Update` of /root/reference/python_deprecated/gravSolver.py:91-118 run
    in the build container (tests/golden/make_legacy_golden.py): fixed dt = 0.2, psi after 1 / 10 / 40 steps.  The
    oracle takes the same dt when the dump interval is the binding constraint (cfl huge, final_time / dumps = dt)."""
    z = np.load(os.path.join(GOLDEN, "legacy_grav3d_traj.npz"))
    n, length, hbar_, mtot, c, dt = z["params"]
    assert c == o.POIS_CONST
    p = o.SimulationParameters(axis_length=float(length), time=0.0, final_sim_time=float(dt) * 40, cfl=1e9, num_data_dumps=40,
                               total_mass=float(mtot), particle_mass=o.HBAR / float(hbar_), sim_name="legacy",
                               k2_cutoff=0.95, alias_threshold=1e9, hbar_=float(hbar_), dims=3, size=int(n))
    return z, p, float(dt)


def test_integrator_step_matches_the_reference_legacy_python_run():
    """Pins drift, density, Poisson solve (constant, sign, k = 0 -> 0), kick and their order in `SimulationObject.update`
    (restating simulation_object.rs:504-581, :1031-1110) to code written by the reference's authors and EXECUTED here:
    1 step 2.5e-16, 40 steps of a collapsing Gaussian (kick phase 0.45 rad per step) 5e-14."""
    z, p, dt = legacy_parameters()
    sim = o.SimulationObject(p, z["psi0"])
    sim.calculate_potential()
    assert rel_l2(sim.phi.real, z["phi0"]) < 1e-14
    for k in range(1, 41):
        sim.update()
        assert abs(sim.last_dt - dt) <= 1e-15                      # dump-limited: exactly the legacy run's fixed dt
        if k in (1, 10, 40):
            assert rel_l2(sim.psi, z[f"psi_{k:03d}"]) < (1e-14 if k == 1 else 1e-12), k
