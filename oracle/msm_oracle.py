"""CPU oracle for the MSM `simulator` time-evolution loop.  TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy/SciPy (pocketfft, fp64) restatement of the reference's
split-step Schroedinger-Poisson integrator.  It is NOT part of the product: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and there only as the checker / the CPU baseline.  The product
path (`msm_b200`) never imports anything from `oracle/`.

Every function cites the reference file:line (relative to the MSM repository root) that it
follows.  The operation order of the reference is kept on purpose (7 complex 3-D FFTs per
step, potential solved twice, phi kept complex until `real()`), so that this file is also the
"un-fused reference step" timed as the CPU baseline.

PINNING STATUS
--------------
* pinned against the reference's own tests (see tests/test_oracle_pins.py):
  `get_kgrid`  (utils/fft.rs:164-167), `spec_grid` (utils/fft.rs:185-214),
  unitary FFT normalisation + invertibility (simulator/tests/fft.rs:2-64 and the other 11
  round trips), `normalize` (utils/grid.rs:107-186), `parse_seeds` (common/src/parameters.rs:121-144),
  TOML parsing of the shipped example files, npz IC fixture loading.
* pinned against the reference's own LEGACY PYTHON integrator, executed in the build container
  (tests/golden/make_legacy_golden.py -> tests/golden/legacy_grav3d_traj.npz): `Grav3D.Update` +
  `compute_phi` of python_deprecated/gravSolver.py:77-118 is the authors' earlier implementation of
  the same static-box split step (half drift, density, Poisson solve with phi_k[0] = 0, full kick,
  half drift) with a fixed dt.  `SimulationObject.update` / `calculate_potential` agree with that
  run to 2.5e-16 after 1 step and 5e-14 after 40 steps of a collapsing Gaussian (kick phase
  0.45 rad per step), phi to 1.5e-16 (test_integrator_step_matches_the_reference_legacy_python_run).
  This pins the drift / kick phases, the density prefactor, the Poisson constant, sign and DC
  handling, the transform conventions and the order of operations of the static-box step.
  (The legacy tree also corroborates the sampler FORMULAS -- python_deprecated/testSinWaveCollapse.py:113-129 adds
  sqrt(.5/n) (N(0, 1/2) + i N(0, 1/2)) for Wigner and sqrt(2) times that for Husimi, i.e. the (N + iN) / (2 sqrt n)
  and (N + iN) / (sqrt 2 sqrt n) of ics.rs:578-585 / :622-629 -- but not a random stream.)
* PARITY UNPINNED for: the adaptive time step and dump bookkeeping (`get_timestep`), `check_alias`
  (the legacy code normalises differently), the expanding-box step, the sampler's random stream
  (ArrayFire Philox -> normal is not reproducible without ArrayFire; the Poisson scheme is unseeded
  in the reference, ics.rs:497) and the scale factor a(t) (crate `cosmology` 0.2.0 is not
  vendored).  The reference holds no golden data for any of these (its Rust tests never call
  `update`), and the Rust crate cannot be built here (no cargo/rustc, ArrayFire 3.8.0 binary is
  downloaded by simulator/build.rs:10-11).  For these the oracle is a line-by-line restatement.

Third-party code the reference's arithmetic lives in (un-vendored):
  arrayfire crate =3.8.0 + ArrayFire v3.8.0 binary (simulator/Cargo.toml:14, build.rs:10-11)
  cosmology crate 0.2.0 (simulator/Cargo.toml:29) -- restated here as an RK4 integration of the
  flat-FLRW Friedmann equation with sub-steps bounded by max_dloga.
  rand 0.8.5 / rand_distr 0.4.3 (Poisson sampler, ics.rs:497,519).

Array layout: the reference's linear buffers are dim-0-fastest (ArrayFire); NPY I/O reshapes the
same buffer C-order (io.rs:63-66), so NumPy index [i][j][k] <-> linear i*n*n + j*n + k and the
LAST NumPy axis is ArrayFire's dim 0.  Every operator on this path is symmetric under axis
permutation, so this only matters for I/O and for the summation order inside `spec_grid`.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field, replace
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft as _fft

# --------------------------------------------------------------------------------------
# constants  (common/src/constants.rs:2-9)
# --------------------------------------------------------------------------------------
POIS_CONST = 4.0 * math.pi * 4.49e-12      # constants.rs:2
HBAR = 1.757e-90                           # constants.rs:5
LITTLE_H_TO_BIG_H = 1.022e-4               # constants.rs:9
DEFAULT_MAX_DLOGA = 1e-3                   # expanding.rs:27

_WORKERS = int(os.environ.get("MSM_ORACLE_WORKERS", os.cpu_count() or 1))


def set_workers(n: int) -> None:
    """Number of pocketfft worker threads (the CPU baseline states this as `cores`)."""
    global _WORKERS
    _WORKERS = max(1, int(n))


def get_workers() -> int:
    return _WORKERS


_POOL = None
_PAR_MIN = 1 << 20      # grids below 2^20 cells are evaluated in one piece (bit-identical to the plain expression)


def _pmap(fn, *arrays):
    """fn(*arrays) evaluated slab by slab (axis 0) on the worker threads -- NumPy releases the GIL inside ufuncs, so the
    element-wise passes of the step use all host cores like the FFTs do.  Element-wise results are bit-identical."""
    global _POOL
    a0 = arrays[0]
    if _WORKERS <= 1 or a0.size < _PAR_MIN or a0.shape[0] < 2:
        return fn(*arrays)
    from concurrent.futures import ThreadPoolExecutor
    if _POOL is None or _POOL._max_workers != _WORKERS:
        _POOL = ThreadPoolExecutor(max_workers=_WORKERS)
    nchunk = min(a0.shape[0], 4 * _WORKERS)
    bounds = np.linspace(0, a0.shape[0], nchunk + 1).astype(int)
    parts = list(_POOL.map(lambda i: fn(*(a[bounds[i]:bounds[i + 1]] for a in arrays)), range(nchunk)))
    if np.ndim(parts[0]) == 0:
        return parts                      # per-slab scalars (reductions): the caller combines them
    return np.concatenate(parts, axis=0)


# --------------------------------------------------------------------------------------
# FFT wrappers + spectral grid  (simulator/src/utils/fft.rs)
# --------------------------------------------------------------------------------------
def forward(a: np.ndarray) -> np.ndarray:
    """fft.rs:6-31 `forward`: d-dim C2C DFT, e^{-i...}, scaled by size^{-dims/2} (unitary)."""
    return _fft.fftn(a, norm="ortho", workers=_WORKERS)


def inverse(a: np.ndarray) -> np.ndarray:
    """fft.rs:33-58 `inverse`: d-dim inverse C2C DFT, also scaled by size^{-dims/2}."""
    return _fft.ifftn(a, norm="ortho", workers=_WORKERS)


def get_kgrid(dx: float, size: int) -> np.ndarray:
    """fft.rs:100-120 `get_kgrid`: k_i = i/(n dx) for i<n/2 else (i-n)/(n dx); n must be even."""
    assert size % 2 == 0                                   # fft.rs:105
    i = np.arange(size, dtype=np.int64)
    i = np.where(i < size // 2, i, i - size)
    return i.astype(np.float64) / (float(size) * dx)


def spec_grid(dx: float, dims: int, size: int) -> np.ndarray:
    """fft.rs:123-161 `spec_grid`: k^2 = (2 pi)^2 * sum_axes k_i^2.

    Summation order follows the reference: zeros + k^2(dim0) + k^2(dim1) + k^2(dim2), then the
    (2 pi)^2 factor.  ArrayFire dim 0 is NumPy's last axis."""
    k2 = get_kgrid(dx, size) ** 2
    arr = np.zeros((size,) * dims, dtype=np.float64)
    for d in range(dims):                                  # fft.rs:141-152
        shape = [1] * dims
        shape[dims - 1 - d] = size
        arr = arr + k2.reshape(shape)
    return arr * ((2.0 * math.pi) ** 2.0)                  # fft.rs:154-160


# --------------------------------------------------------------------------------------
# grid helpers  (simulator/src/utils/grid.rs)
# --------------------------------------------------------------------------------------
def normalize(grid: np.ndarray, dx: float, dims: int) -> np.ndarray:
    """grid.rs:11-33 `normalize`: scale so that sum|grid|^2 * dx^dims == 1."""
    norm = np.sum((grid * np.conj(grid)).real)
    return grid * math.sqrt(dx ** (-float(dims)) / norm)


def check_norm(grid: np.ndarray, dx: float, dims: int) -> bool:
    """grid.rs:35-64 `check_norm` (tolerance 1e-4)."""
    norm = float(np.sum((grid * np.conj(grid)).real))
    return abs(norm * dx ** float(dims) - 1.0) < 1e-4


# --------------------------------------------------------------------------------------
# configuration  (common/src/parameters.rs, common/src/ics.rs)
# --------------------------------------------------------------------------------------
@dataclass
class CosmologyParameters:
    """parameters.rs:71-86."""
    omega_matter_now: float
    omega_radiation_now: float
    h: float
    z0: float
    max_dloga: Optional[float] = None


@dataclass
class TomlParameters:
    """parameters.rs:11-55 (remote storage table ignored: out of scope)."""
    axis_length: float
    final_sim_time: float
    cfl: float
    num_data_dumps: int
    total_mass: float
    sim_name: str
    k2_cutoff: float
    alias_threshold: float
    dims: int
    size: int
    ics: dict
    time: Optional[float] = None
    particle_mass: Optional[float] = None
    ntot: Optional[float] = None
    hbar_: Optional[float] = None
    sampling: Optional[dict] = None        # {"scheme": str, "seeds": [int]}
    output_potential: bool = False
    cosmology: Optional[CosmologyParameters] = None


def parse_seeds(s: str) -> List[int]:
    """parameters.rs:148-202 `parse_seeds` ("a..=b", "a to b", "[s1, s2]", "s1, s2")."""
    if re.search(r"\d+..=\d+", s):
        a, b = (int(x) for x in s.split("..="))
        return list(range(a, b + 1))
    if re.search(r"\d+ to \d+", s):
        a, b = (int(x) for x in s.split(" to "))
        return list(range(a, b + 1))
    found = re.findall(r"(\d+)[^,]?", s)
    if found:
        return [int(x) for x in found]
    raise ValueError("seeds did not match expected patterns: low..=high, low to high, [s1, s2, s3]")


def read_toml(path: str) -> TomlParameters:
    """parameters.rs:96-107 `read_toml` (serde ignores unknown keys such as `num_streams`)."""
    import tomllib
    with open(path, "rb") as f:
        d = tomllib.load(f)
    sampling = None
    if "sampling" in d:
        sampling = {"scheme": d["sampling"]["scheme"], "seeds": parse_seeds(d["sampling"]["seeds"])}
    cosmo = None
    if "cosmology" in d:
        c = d["cosmology"]
        cosmo = CosmologyParameters(float(c["omega_matter_now"]), float(c["omega_radiation_now"]),
                                    float(c["h"]), float(c["z0"]),
                                    float(c["max_dloga"]) if "max_dloga" in c else None)
    opt = lambda k: float(d[k]) if k in d else None
    return TomlParameters(
        axis_length=float(d["axis_length"]), final_sim_time=float(d["final_sim_time"]), cfl=float(d["cfl"]),
        num_data_dumps=int(d["num_data_dumps"]), total_mass=float(d["total_mass"]), sim_name=str(d["sim_name"]),
        k2_cutoff=float(d["k2_cutoff"]), alias_threshold=float(d["alias_threshold"]), dims=int(d["dims"]),
        size=int(d["size"]), ics=dict(d["ics"]), time=opt("time"), particle_mass=opt("particle_mass"),
        ntot=opt("ntot"), hbar_=opt("hbar_"), sampling=sampling,
        output_potential=bool(d.get("output_potential", False)), cosmology=cosmo)


def determine_pmass_hbar_(toml: TomlParameters) -> Tuple[float, float]:
    """parameters.rs:222-259 `determine_pmass_hbar_`."""
    if toml.ntot is not None:
        particle_mass = toml.total_mass / toml.ntot
        hbar_ = toml.hbar_ if toml.hbar_ is not None else HBAR / particle_mass
    elif toml.particle_mass is not None:
        particle_mass = toml.particle_mass
        hbar_ = toml.hbar_ if toml.hbar_ is not None else HBAR / particle_mass
    elif toml.hbar_ is not None:
        hbar_ = toml.hbar_
        particle_mass = HBAR / hbar_
    else:
        raise ValueError("You must specify the total mass and one of ntot, particle_mass or hbar_")
    return particle_mass, hbar_


def get_supercomoving_boxsize(hbar_: float, cosmo: CosmologyParameters, axis_length: float) -> float:
    """parameters.rs:205-220 `get_supercomoving_boxsize`."""
    initial_scale_factor = 1.0 / (1.0 + cosmo.z0)
    comoving_boxsize = axis_length / initial_scale_factor
    return math.sqrt(math.sqrt(1.5 * cosmo.omega_matter_now * (LITTLE_H_TO_BIG_H * cosmo.h) ** 2) / hbar_) \
        * comoving_boxsize


# --------------------------------------------------------------------------------------
# cosmology scalars  (simulator/src/expanding.rs, utils/mod.rs, simulation_object.rs:1344-1453)
# --------------------------------------------------------------------------------------
def rk4(f, tn: float, yn: float, h: float, derivative: Optional[float] = None) -> float:
    """utils/mod.rs:14-43 `rk4`: one classical RK4 step of y' = f(t, y)."""
    k1 = derivative if derivative is not None else f(tn, yn)
    k2 = f(tn + h / 2.0, yn + h * k1 / 2.0)
    k3 = f(tn + h / 2.0, yn + h * k2 / 2.0)
    k4 = f(tn + h, yn + h * k3)
    return yn + h * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0


class ScaleFactorSolver:
    """expanding.rs:12-118 `ScaleFactorSolver`.

    The reference delegates to crate `cosmology` 0.2.0 (`scale_factor::ScaleFactor`), whose source
    is not vendored.  Restated as: flat FLRW, omega_de0 = 1 - omega_m0 - omega_r0 (expanding.rs:29-38),
    H0 = h * 1.022e-4 / Myr, a(t0 = 0) = 1/(1+z0), da/dt = a H0 sqrt(Om a^-3 + Or a^-4 + Ode),
    integrated by RK4 sub-steps with |dt_sub| <= max_dloga * a / (da/dt).  Negative `dt` steps
    backwards with the same rule (needed by `calculate_dt_from_dtau`, whose RK4 stages may ask for a
    slightly earlier time than the previous stage)."""

    def __init__(self, cosmo: CosmologyParameters):
        assert cosmo.omega_matter_now + cosmo.omega_radiation_now <= 1.0      # expanding.rs:62-65
        assert cosmo.z0 >= 0.0 and cosmo.omega_matter_now >= 0.0 and cosmo.omega_radiation_now >= 0.0
        self.cosmo = cosmo
        self.om = cosmo.omega_matter_now
        self.orad = cosmo.omega_radiation_now
        self.ode = 1.0 - self.om - self.orad
        self.h0 = cosmo.h * LITTLE_H_TO_BIG_H
        self.max_dloga = cosmo.max_dloga if cosmo.max_dloga is not None else DEFAULT_MAX_DLOGA
        self.a = 1.0 / (1.0 + cosmo.z0)
        self.t = 0.0                                                            # expanding.rs:82

    def clone(self) -> "ScaleFactorSolver":
        c = ScaleFactorSolver.__new__(ScaleFactorSolver)
        c.__dict__.update(self.__dict__)
        return c

    def _dadt(self, a: float) -> float:
        return a * self.h0 * math.sqrt(self.om / a ** 3 + self.orad / a ** 4 + self.ode)

    def step(self, dt: float) -> float:
        """expanding.rs:99-105 `step` -> inner `step_forward(dt)`; returns the new a."""
        remaining = dt
        while remaining != 0.0:
            lim = self.max_dloga * self.a / self._dadt(self.a)
            h = remaining if abs(remaining) <= lim else math.copysign(lim, remaining)
            a = self.a
            k1 = self._dadt(a)
            k2 = self._dadt(a + 0.5 * h * k1)
            k3 = self._dadt(a + 0.5 * h * k2)
            k4 = self._dadt(a + h * k3)
            self.a = a + h * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
            self.t += h
            remaining = 0.0 if h == remaining else remaining - h
        return self.a

    def get_a(self) -> float:          # expanding.rs:107-109
        return self.a

    def get_dadt(self) -> float:       # expanding.rs:111-113
        return self._dadt(self.a)

    def get_time(self) -> float:       # expanding.rs:115-117
        return self.t


def get_tau(target_time: float, cosmo: CosmologyParameters) -> float:
    """simulation_object.rs:1408-1453 `get_tau`: integrate dtau/dt = sqrt(1.5 Om H0^2)/a^2 from t=0
    with a FRESH solver, steps dt = min(target/1000, a/(da/dt)*max_dloga, target - time)."""
    solver = ScaleFactorSolver(cosmo)
    pref = math.sqrt(1.5 * cosmo.omega_matter_now * (LITTLE_H_TO_BIG_H * cosmo.h) ** 2)

    def dtau_dt(t: float, _tau: float) -> float:
        a_at_t = solver.step(t - solver.get_time())                            # :1420-1423
        return pref / a_at_t ** 2                                               # :1426-1428

    tau = 0.0
    time = 0.0
    while time < target_time:                                                   # :1436
        dt = target_time / 1000.0
        if cosmo.max_dloga is not None:                                         # :1438-1443
            dt = min(target_time / 1000.0, solver.get_a() / solver.get_dadt() * cosmo.max_dloga)
        dt = min(dt, target_time - time)                                        # :1444
        tau = rk4(dtau_dt, time, tau, dt)                                       # :1447
        time += dt
    return tau


# --------------------------------------------------------------------------------------
# SimulationParameters  (simulation_object.rs:67-140, ::new :223-315)
# --------------------------------------------------------------------------------------
@dataclass
class SimulationParameters:
    axis_length: float
    time: float
    final_sim_time: float
    cfl: float
    num_data_dumps: int
    total_mass: float
    particle_mass: float
    sim_name: str
    k2_cutoff: float
    alias_threshold: float
    hbar_: float
    dims: int
    size: int
    output_potential: bool = False
    cosmo_params: Optional[CosmologyParameters] = None      # None <=> static box (feature off)
    sampling_parameters: Optional[dict] = None               # {"seed": int, "scheme": str}
    ics: Optional[dict] = None
    # derived (filled by __post_init__, simulation_object.rs:243-274)
    dx: float = 0.0
    dk: float = 0.0
    n_tot: float = 0.0
    k2_max: float = 0.0
    comoving_boxsize: float = 0.0
    tau: float = 0.0
    final_sim_tau: float = 0.0
    current_dumps: int = 0
    n_steps: int = 0
    spec_grid: np.ndarray = field(default=None, repr=False)

    @property
    def expanding(self) -> bool:
        return self.cosmo_params is not None

    def __post_init__(self):
        if self.expanding:
            self.tau = get_tau(self.time, self.cosmo_params)                    # :246
            self.final_sim_tau = get_tau(self.final_sim_time, self.cosmo_params)  # :248-249
            self.comoving_boxsize = get_supercomoving_boxsize(self.hbar_, self.cosmo_params, self.axis_length)
            self.dx = self.comoving_boxsize / float(self.size)                  # :262
        else:
            self.dx = self.axis_length / float(self.size)                       # :260
        self.dk = self.dx                                                       # :263 (sic)
        self.n_tot = self.total_mass / self.particle_mass                       # :264
        self.spec_grid = spec_grid(self.dx, self.dims, self.size)               # :273
        self.k2_max = float(self.spec_grid.max())                               # :274


def simulation_iter(toml: TomlParameters, expanding: Optional[bool] = None) -> Iterator[SimulationParameters]:
    """utils/io.rs:127-245 `parameters_from_toml` + `SimulationIter::next`: one parameter set per seed,
    named "<sim>-stream%05d", then ONE un-sampled mean-field run named "<sim>".

    `expanding` mirrors the compile-time cargo feature; default: on iff the TOML has [cosmology]."""
    particle_mass, hbar_ = determine_pmass_hbar_(toml)                          # io.rs:166
    if expanding is None:
        expanding = toml.cosmology is not None
    cosmo = toml.cosmology if expanding else None
    seeds = list(toml.sampling["seeds"]) if toml.sampling else []
    common = dict(axis_length=toml.axis_length, time=toml.time if toml.time is not None else 0.0,
                  final_sim_time=toml.final_sim_time, cfl=toml.cfl, num_data_dumps=toml.num_data_dumps,
                  total_mass=toml.total_mass, particle_mass=particle_mass, k2_cutoff=toml.k2_cutoff,
                  alias_threshold=toml.alias_threshold, hbar_=hbar_, dims=toml.dims, size=toml.size,
                  output_potential=toml.output_potential, cosmo_params=cosmo, ics=dict(toml.ics))
    for seed in seeds:                                                          # io.rs:183-213
        yield SimulationParameters(sim_name=f"{toml.sim_name}-stream{seed:05d}",
                                   sampling_parameters={"seed": seed, "scheme": toml.sampling["scheme"]}, **common)
    yield SimulationParameters(sim_name=toml.sim_name, sampling_parameters=None, **common)   # io.rs:214-240


# --------------------------------------------------------------------------------------
# counter-based RNG used by the restated sampler (NOT ArrayFire's stream: unpinned, see header)
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox-4x32-10 (Salmon et al. 2011), vectorised over counter arrays (uint64 holding 32-bit words)."""
    c0 = np.asarray(c0, dtype=np.uint64) & _M32
    c1 = np.asarray(c1, dtype=np.uint64) & _M32
    c2 = np.asarray(c2, dtype=np.uint64) & _M32
    c3 = np.asarray(c3, dtype=np.uint64) & _M32
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(hi, lo):
    """two 32-bit words -> uniform in (0,1) with 53 random bits, never 0 or 1."""
    return ((hi >> np.uint64(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint64(6)).astype(np.float64) + 0.5) \
        * (1.0 / 9007199254740992.0)


def philox_uniform(n_cells: int, seed: int, draw: int = 0, start: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Two uniforms per cell: counter = (cell_lo, cell_hi, draw, 0), key = (seed_lo, seed_hi); cells start .. start +
    n_cells - 1 (the stream is counter-based, so a grid can be generated slab by slab)."""
    q = np.arange(start, start + n_cells, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(q & _M32, q >> np.uint64(32), np.full(n_cells, draw, np.uint64),
                                   np.zeros(n_cells, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return _u53(x0, x1), _u53(x2, x3)


def philox_normal_pair(n_cells: int, seed: int, draw: int = 0, start: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Box-Muller on `philox_uniform`: two independent N(0,1) per cell."""
    u1, u2 = philox_uniform(n_cells, seed, draw, start)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * math.pi * u2
    return r * np.cos(ang), r * np.sin(ang)


# --------------------------------------------------------------------------------------
# initial conditions  (simulator/src/ics.rs)
# --------------------------------------------------------------------------------------
def cold_gauss(mean: Sequence[float], std: Sequence[float], p: SimulationParameters) -> np.ndarray:
    """ics.rs:24-162 `cold_gauss`: separable real Gaussian, x_i = (2i+1) dx/2, normalised."""
    assert len(mean) == p.dims and len(std) == p.dims                          # :60-69
    x = (2.0 * np.arange(p.size) + 1.0) * p.dx / 2.0                            # :72-74
    comps = []
    for d in range(p.dims):
        g = np.exp(-0.5 * ((x - mean[d]) / std[d]) ** 2.0).astype(np.complex128)  # :79-87
        comps.append(normalize(g, p.dx, p.dims))                                # :91 (dx^dims on a 1-D array, sic)
    # ics.rs:140-141: psi = psi_x (AF dim 0) * psi_y (dim 1) * psi_z (dim 2); AF dim 0 = last NumPy axis
    psi = comps[0]
    if p.dims >= 2:
        psi = comps[1][:, None] * psi[None, :]
    if p.dims == 3:
        psi = comps[2][:, None, None] * psi[None, :, :]
    return normalize(psi, p.dx, p.dims)                                         # :142


def spherical_tophat(p: SimulationParameters, radius: float, delta: float, slope: float) -> np.ndarray:
    """ics.rs:165-280 `spherical_tophat`: sqrt(1 + delta * ramp(r)), ramp = 1/(1+exp(slope (r/R - 1)))."""
    dx = p.axis_length / float(p.size)                                          # :203
    x = (2.0 * np.arange(p.size) + 1.0) * dx / 2.0                              # :207-209
    half = p.axis_length / 2.0
    xi = x - half
    if p.dims == 3:
        r = np.sqrt(xi[:, None, None] ** 2 + xi[None, :, None] ** 2 + xi[None, None, :] ** 2)
    elif p.dims == 2:
        # z collapses to the single value `null` = L/2 -> zk = 0  (:206-219)
        r = np.sqrt(xi[:, None] ** 2 + xi[None, :] ** 2 + 0.0)
    else:
        r = np.sqrt(xi ** 2 + 0.0 + 0.0)
    ramp = 1.0 / (1.0 + np.exp(slope * (r / radius - 1.0)))                     # :221-226
    psi = np.sqrt(1.0 + delta * ramp).astype(np.complex128)                     # :238-241
    return normalize(psi, p.dx, p.dims)                                         # :261


def cold_gauss_kspace(mean, std, p: SimulationParameters, seed: Optional[int]) -> np.ndarray:
    """ics.rs:282-431 `cold_gauss_kspace`: Gaussian in k with uniform random phases, then forward FFT.
    The uniform phases come from ArrayFire's Philox in the reference (:400,:419); here from `philox_uniform`
    (draw slot 7) -- statistically equivalent, not bit-identical (unpinned).  The reference hard-codes a 3-D
    phase array for any `dims` (:401-406); only dims == 3 is meaningful there and only that is restated."""
    assert len(mean) == p.dims and len(std) == p.dims
    k = get_kgrid(p.dx, p.size)                                                 # :331
    comps = []
    for d in range(p.dims):
        g = np.exp(-0.5 * ((k - mean[d]) / std[d]) ** 2.0).astype(np.complex128)
        comps.append(normalize(g, p.dk, p.dims))                                # :347
    psi = comps[0]
    if p.dims >= 2:
        psi = comps[1][:, None] * psi[None, :]
    if p.dims == 3:
        psi = comps[2][:, None, None] * psi[None, :, :]
    psi = normalize(psi, p.dk, p.dims)                                          # :395
    u, _ = philox_uniform(psi.size, 0 if seed is None else int(seed), draw=7)  # :399
    psi = psi * np.exp(2j * math.pi * u.reshape(psi.shape))                     # :407-423
    return forward(psi)                                                         # :425


def user_specified_ics(path: str, p: SimulationParameters) -> np.ndarray:
    """ics.rs:650-730 `user_specified_ics`: npz with members real.npy / imag.npy, not renormalised."""
    z = np.load(path)
    re_, im_ = np.asarray(z["real"], dtype=np.float64), np.asarray(z["imag"], dtype=np.float64)
    if re_.ndim != p.dims:
        raise ValueError("Dimensions of user-provided data do not match the dimensions specified in the toml")
    if any(s != re_.shape[0] for s in re_.shape):
        raise ValueError("Only uniform grids are supported at this time")
    if re_.shape[0] != p.size:
        raise ValueError("Grid size of user-provided data does not match the size specified in the toml")
    return re_ + 1j * im_


def sample_quantum_perturbation(psi: np.ndarray, p: SimulationParameters, sampling: dict) -> np.ndarray:
    """ics.rs:434-648 `sample_quantum_perturbation` (Poisson :495-558, Wigner :560-602, Husimi :604-646).
    Normals: `philox_normal_pair(cell, seed)` (draw slot 0) instead of ArrayFire's stream (unpinned)."""
    n = p.total_mass / p.particle_mass                                          # :472
    sqrt_n = math.sqrt(n)
    sqrt_dv = math.sqrt(p.dx ** float(p.dims))                                  # :481-486
    seed, scheme = int(sampling["seed"]), sampling["scheme"]
    psi_count = psi * sqrt_dv                                                   # :479-489
    if scheme == "Poisson":
        # reference uses the unseeded thread_rng (:497); here seeded for reproducibility
        rng = np.random.Generator(np.random.Philox(seed))
        lam = np.abs((psi * np.conj(psi))) * p.dx ** float(p.dims) * n          # :509-515 (uses psi, not psi_count)
        a = rng.poisson(lam).astype(np.float64)
        mag = np.sqrt(a / n)                                                    # :523
        out = mag * np.exp(1j * np.angle(psi))                                  # :535-544
        return out / sqrt_dv                                                    # :547-557
    if psi.size >= _PAR_MIN and _WORKERS > 1:     # slab by slab on the worker threads: the same numbers, less memory
        def slab(a):
            lo = int(a[0].reshape(-1)[0])
            z0, z1 = philox_normal_pair(a.size, seed, draw=0, start=lo)
            return (z0 + 1j * z1).reshape(a.shape)
        samples = _pmap(slab, np.arange(psi.size, dtype=np.int64).reshape(psi.shape))
    else:
        z0, z1 = philox_normal_pair(psi.size, seed, draw=0)
        samples = (z0 + 1j * z1).reshape(psi.shape)                             # :563-575 / :607-619
    if scheme == "Wigner":
        samples = samples / (sqrt_n * 2.0)                                      # :578-585
    elif scheme == "Husimi":
        samples = samples / (sqrt_n * math.sqrt(2.0))                           # :622-629
    else:
        raise ValueError(f"unknown sampling scheme {scheme}")
    return (psi_count + samples) / sqrt_dv                                      # :588-601 / :632-645


def initial_wavefunction(p: SimulationParameters, base_dir: str = ".") -> np.ndarray:
    """simulation_object.rs:404-435 `new_from_params`: build the IC, then apply the sampler if any."""
    ics = p.ics
    t = ics["type"]
    if t == "UserSpecified":
        psi = user_specified_ics(os.path.join(base_dir, ics["path"]), p)
    elif t == "ColdGauss":
        psi = cold_gauss([float(v) for v in ics["mean"]], [float(v) for v in ics["std"]], p)
    elif t == "ColdGaussKSpace":
        psi = cold_gauss_kspace([float(v) for v in ics["mean"]], [float(v) for v in ics["std"]], p,
                                ics.get("phase_seed"))
    elif t == "SphericalTophat":
        psi = spherical_tophat(p, float(ics["radius"]), float(ics["delta"]), float(ics["slope"]))
    else:
        raise ValueError(f"unknown ics type {t}")
    if p.sampling_parameters is not None:
        psi = sample_quantum_perturbation(psi, p, p.sampling_parameters)
    return np.ascontiguousarray(psi, dtype=np.complex128)


# --------------------------------------------------------------------------------------
# the integrator  (simulator/src/simulation_object.rs)
# --------------------------------------------------------------------------------------
class FourierAliasing(RuntimeError):
    """utils/error.rs:5-27 `RuntimeError::FourierAliasing` (the reference panics, :607-617)."""

    def __init__(self, threshold, k2_cutoff, p_mass):
        super().__init__(f"simulation aliased: threshold {threshold} k2_cutoff {k2_cutoff} p_mass {p_mass}")
        self.threshold, self.k2_cutoff, self.p_mass = threshold, k2_cutoff, p_mass


class SimulationObject:
    """simulation_object.rs:145-184 `SimulationObject` with grid {psi, psik, phi} (:42-64)."""

    def __init__(self, parameters: SimulationParameters, psi0: np.ndarray):
        self.parameters = parameters
        self.psi = np.ascontiguousarray(psi0, dtype=np.complex128)
        self.psik = self.psi.copy()                 # :206 "initialised with incorrect values"
        self.phi = self.psi.real.astype(np.complex128)   # :205
        self.dumps: List[Tuple[int, np.ndarray]] = []
        self.last_dt = 0.0
        self.last_alias_mass = 0.0
        self.last_potential_max = 0.0
        self.scale_factor_solver = ScaleFactorSolver(parameters.cosmo_params) if parameters.expanding else None

    # ---- density / potential ---------------------------------------------------------
    def density_prefactor(self) -> float:
        """simulation_object.rs:1033-1056: static: total_mass; expanding: M C (2/(3 H0^2 Om))^(1/4) / hbar_^(d/2)."""
        p = self.parameters
        if p.expanding:
            c = p.cosmo_params
            return p.total_mass * POIS_CONST * (2.0 / (3.0 * (c.h * LITTLE_H_TO_BIG_H) ** 2 * c.omega_matter_now)) \
                ** (1.0 / 4.0) / p.hbar_ ** (float(p.dims) / 2.0)
        return p.total_mass

    def poisson_coeff(self) -> float:
        """simulation_object.rs:1079-1086: expanding -1, static -POIS_CONST."""
        return -1.0 if self.parameters.expanding else -POIS_CONST

    def calculate_density(self) -> None:
        """simulation_object.rs:1031-1063: phi <- A * real(psi conj(psi)) cast to complex."""
        A = self.density_prefactor()
        self.phi = _pmap(lambda a: (A * (a * np.conj(a)).real).astype(np.complex128), self.psi)

    def calculate_potential(self) -> None:
        """simulation_object.rs:1066-1110: phi = Re F^-1[ c F[rho] / k^2 , k=0 -> 0 ]."""
        p = self.parameters
        self.calculate_density()                                                # :1069
        self.phi = forward(self.phi)                                            # :1071
        c = complex(self.poisson_coeff(), 0.0)

        def solve(f, k2):
            with np.errstate(divide="ignore", invalid="ignore"):
                out = (c * f) / k2.astype(np.complex128)                        # :1076-1095
            out[np.isnan(out)] = 0.0                                            # :1098-1102 (0/0 at k = 0)
            return out

        self.phi = _pmap(solve, self.phi, p.spec_grid)
        self.phi = inverse(self.phi)                                            # :1105
        self.phi = _pmap(lambda f: f.real.astype(np.complex128), self.phi)      # :1109

    # ---- time step -------------------------------------------------------------------
    def get_timestep(self) -> Tuple[bool, float]:
        """simulation_object.rs:878-934 (static) / :939-990 (expanding)."""
        p = self.parameters
        potential_max = float(np.max(_pmap(lambda f: np.max(np.abs(f)), self.phi)))   # :905 / :954
        self.last_potential_max = potential_max
        time_to_next_dump = (float(p.current_dumps + 1) * p.final_sim_time / float(p.num_data_dumps)) - p.time
        if not p.expanding:
            kinetic_dt = p.cfl * 2.0 * p.axis_length / math.sqrt(p.k2_max) / p.hbar_           # :881-884
            potential_dt = p.cfl * (2.0 * math.pi) * p.hbar_ / (2.0 * potential_max)           # :906-909
            dt = min(min(kinetic_dt, potential_dt), time_to_next_dump)                          # :922
            return dt == time_to_next_dump, dt                                                  # :926-933
        kinetic_dtau = p.cfl * 2.0 * p.comoving_boxsize / math.sqrt(p.k2_max)                   # :942-944
        potential_dtau = p.cfl * (2.0 * math.pi) / ((2.0 * self.scale_factor_solver.get_a()) * potential_max)  # :957-959
        tau_to_next_dump = get_tau(p.time + time_to_next_dump, p.cosmo_params) - p.tau          # :970-975
        dtau = min(min(kinetic_dtau, potential_dtau), tau_to_next_dump)                         # :978
        return dtau == tau_to_next_dump, dtau                                                   # :982-989

    def calculate_dt_from_dtau(self, dtau: float) -> float:
        """simulation_object.rs:1344-1388: one RK4 step of dt/dtau = a(t)^2 / sqrt(1.5 Om H0^2) on a CLONED solver."""
        p = self.parameters
        solver = self.scale_factor_solver.clone()                               # :1347
        pref = math.sqrt(1.5 * p.cosmo_params.omega_matter_now * (LITTLE_H_TO_BIG_H * p.cosmo_params.h) ** 2)

        def dt_dtau(_tau: float, t: float) -> float:
            a_at_t = solver.step(t - solver.get_time())                         # :1352-1355
            return 1.0 / (pref / a_at_t ** 2)                                   # :1358-1363

        return rk4(dt_dtau, p.tau, p.time, dtau) - p.time                       # :1366-1374

    # ---- alias -----------------------------------------------------------------------
    def check_alias(self) -> Optional[float]:
        """simulation_object.rs:1249-1293: p = sum_{k^2 > k2_cutoff k2_max} |psik|^2 dk^dims."""
        p = self.parameters
        cut = p.k2_max * p.k2_cutoff

        def masked_sum(pk, k2):
            a = (pk * np.conj(pk)).real                                         # :1259
            return np.sum(np.where(k2 > cut, a, 0.0))                           # :1262-1280

        p_mass = float(np.sum(_pmap(masked_sum, self.psik, p.spec_grid))) * p.dk ** float(p.dims)   # :1281-1285
        self.last_alias_mass = p_mass
        return p_mass if p_mass > p.alias_threshold else None                   # :1288-1292

    # ---- dump / loop -----------------------------------------------------------------
    def dump(self) -> None:
        """simulation_object.rs:1113-1223: snapshot psi under the current dump index (kept in memory here;
        `write_dump` reproduces the on-disk layout)."""
        self.dumps.append((self.parameters.current_dumps, self.psi.copy()))

    def not_finished(self) -> bool:
        """simulation_object.rs:1226-1228."""
        return self.parameters.time < self.parameters.final_sim_time

    def update(self) -> None:
        """simulation_object.rs:475-661 (static) / :669-873 (expanding)."""
        p = self.parameters
        if p.n_steps == 0:
            self.psik = forward(self.psi)                                       # :477-479 / :672-674
        self.calculate_potential()                                              # :497 / :692
        dump, dt = self.get_timestep()                                          # :500 / :695
        self.last_dt = dt
        kc = complex(0.0, -dt / 4.0 * p.hbar_) if not p.expanding else complex(0.0, -dt / 4.0)   # :504-514 / :699-706
        k_evolution = _pmap(lambda k2: np.exp(kc * k2.astype(np.complex128)), p.spec_grid)
        self.psik = _pmap(np.multiply, self.psik, k_evolution)                  # :516 / :708
        self.psi = inverse(self.psik)                                           # :523 / :715
        self.calculate_potential()                                              # :530 / :722
        if not p.expanding:
            rc = complex(0.0, -dt / p.hbar_)
            r_evolution = _pmap(lambda f: np.exp(rc * f), self.phi)             # :535-542
            self.psi = _pmap(np.multiply, self.psi, r_evolution)                # :545
        else:
            for _ in range(2):                                                  # :726
                a = self.scale_factor_solver.get_a()                            # :728
                rc = complex(0.0, -dt / 2.0 * a)
                r_evolution = _pmap(lambda f: np.exp(rc * f), self.phi)         # :729-739
                self.psi = _pmap(np.multiply, self.psi, r_evolution)            # :742
                dt_half = self.calculate_dt_from_dtau(dt / 2.0)                 # :751-752
                self.scale_factor_solver.step(dt_half)                          # :755-756
                p.time = p.time + dt_half                                       # :757
                p.tau = p.tau + dt / 2.0                                        # :759
        self.psik = forward(self.psi)                                           # :552 / :761
        self.psik = _pmap(np.multiply, self.psik, k_evolution)                  # :562-574 / :771-780 (same array values)
        self.psi = inverse(self.psik)                                           # :581 / :787
        if not p.expanding:
            p.time = p.time + dt                                                # :590
        alias = self.check_alias()                                              # :607 / :815
        if alias is not None:
            raise FourierAliasing(p.alias_threshold, p.k2_cutoff, alias)
        if dump:                                                                # :620-631 / :828-844
            p.current_dumps += 1
            self.dump()
            p.time = float(p.current_dumps) * p.final_sim_time / float(p.num_data_dumps)
            if p.expanding:
                p.tau = get_tau(p.time, p.cosmo_params)
        p.n_steps += 1                                                          # :635 / :797


def run_stream(parameters: SimulationParameters, psi0: np.ndarray, max_steps: Optional[int] = None,
               on_step=None) -> SimulationObject:
    """simulator/src/main.rs:59-69: dump the IC, then `while not_finished() { update() }`."""
    sim = SimulationObject(parameters, psi0)
    sim.dump()                                                                  # main.rs:61
    steps = 0
    while sim.not_finished() and (max_steps is None or steps < max_steps):      # main.rs:65
        sim.update()
        steps += 1
        if on_step is not None:
            on_step(sim)
    return sim


# --------------------------------------------------------------------------------------
# on-disk dump layout  (utils/io.rs:34-108, simulation_object.rs:1155-1158)
# --------------------------------------------------------------------------------------
def dump_shape(dims: int, size: int) -> Tuple[int, int, int, int]:
    """simulation_object.rs:1012-1028 `get_shape_array`."""
    return (size, size if dims >= 2 else 1, size if dims == 3 else 1, 1)


def write_dump(root: str, sim_name: str, dump_index: int, psi: np.ndarray, dims: int, size: int,
               field_name: str = "psi") -> Tuple[str, str]:
    """io.rs:34-88 `complex_array_to_disk`: two extension-less NPY files `<field>_%05d_real|_imag`, f64,
    4-D shape from `dump_shape`, the linear (dim-0-fastest) buffer reshaped C-order (io.rs:63-66)."""
    d = os.path.join(root, sim_name)
    os.makedirs(d, exist_ok=True)
    shape = dump_shape(dims, size)
    flat = np.ascontiguousarray(psi).reshape(-1)
    out = []
    for part, arr in (("real", flat.real), ("imag", flat.imag)):
        path = os.path.join(d, f"{field_name}_{dump_index:05d}_{part}")
        with open(path, "wb") as f:
            np.lib.format.write_array(f, np.ascontiguousarray(arr, dtype=np.float64).reshape(shape), version=(1, 0))
        out.append(path)
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# synthesizer reductions  (synthesizer/src/lib.rs:106-342, synthesizer/src/main.rs:63-93,161-173)
# --------------------------------------------------------------------------------------
def synthesizer_combine(psis: Sequence[np.ndarray], dv: float) -> dict:
    """`analyze_sims`: means over the `-stream*` runs of psi, |psi|^2, psi_k, |psi_k|^2 with psi_k the UN-normalised
    forward DFT (ndrustfft `ndfft` per axis, lib.rs:206-213), and `post_combine`'s Qx = sum(psi2 - |psi|^2) * dv."""
    n = float(len(psis))
    psi = sum(psis) / n                                                         # main.rs:74-77
    psi2 = sum(p * np.conj(p) for p in psis) / n                                # main.rs:78-81
    pk = [_fft.fftn(p, workers=_WORKERS) for p in psis]                         # lib.rs:206-213 (no normalisation)
    psik = sum(pk) / n                                                          # main.rs:82-85
    psik2 = sum(k * np.conj(k) for k in pk) / n                                 # main.rs:86-92
    qx = complex(np.sum(psi2 - psi * np.conj(psi)) * dv)                        # main.rs:161-173
    return {"psi": psi, "psi2": psi2, "psik": psik, "psik2": psik2, "Qx": qx}


# --------------------------------------------------------------------------------------
# the north_star's second coupling mode (NOT in the reference)
# --------------------------------------------------------------------------------------
class SummedEnsemble:
    """S streams that share ONE potential sourced by the ensemble-mean density
    rho = (A / S) * sum_s |psi_s|^2.  Same step sequence as `SimulationObject.update` (static or expanding) with
    a global dt = min over constraints (phi is shared, so all streams get the same dt).  With S identical streams
    (or S == 1) this reproduces the independent-mode trajectory."""

    def __init__(self, parameters: SimulationParameters, psi0s: Sequence[np.ndarray]):
        self.parameters = parameters
        self.streams = [SimulationObject(replace(parameters), p0) for p0 in psi0s]
        self.head = self.streams[0]          # carries time / dump bookkeeping and the a(t) solver

    def _potential(self) -> None:
        h = self.head
        p = h.parameters
        rho = sum((s.psi * np.conj(s.psi)).real for s in self.streams) * (h.density_prefactor() / len(self.streams))
        phik = forward(rho.astype(np.complex128))
        with np.errstate(divide="ignore", invalid="ignore"):
            phik = (complex(h.poisson_coeff(), 0.0) * phik) / p.spec_grid.astype(np.complex128)
        phik[np.isnan(phik)] = 0.0
        phi = inverse(phik).real.astype(np.complex128)
        for s in self.streams:
            s.phi = phi

    def update(self) -> None:
        h = self.head
        p = h.parameters
        if p.n_steps == 0:
            for s in self.streams:
                s.psik = forward(s.psi)
        self._potential()
        dump, dt = h.get_timestep()
        h.last_dt = dt
        coef = -dt / 4.0 * (1.0 if p.expanding else p.hbar_)
        k_evolution = np.exp(complex(0.0, coef) * p.spec_grid.astype(np.complex128))
        for s in self.streams:
            s.psik = s.psik * k_evolution
            s.psi = inverse(s.psik)
        self._potential()
        if not p.expanding:
            r = np.exp(complex(0.0, -dt / p.hbar_) * h.phi)
            for s in self.streams:
                s.psi = s.psi * r
        else:
            for _ in range(2):
                a = h.scale_factor_solver.get_a()
                r = np.exp(complex(0.0, -dt / 2.0 * a) * h.phi)
                for s in self.streams:
                    s.psi = s.psi * r
                dt_half = h.calculate_dt_from_dtau(dt / 2.0)
                h.scale_factor_solver.step(dt_half)
                p.time = p.time + dt_half
                p.tau = p.tau + dt / 2.0
        alias = []
        for s in self.streams:
            s.psik = forward(s.psi) * k_evolution
            s.psi = inverse(s.psik)
            s.parameters.k2_max = p.k2_max
            alias.append(s.check_alias())
        if not p.expanding:
            p.time = p.time + dt
        for i, a in enumerate(alias):
            if a is not None:
                raise FourierAliasing(p.alias_threshold, p.k2_cutoff, a)
        if dump:
            p.current_dumps += 1
            for s in self.streams:
                s.parameters.current_dumps = p.current_dumps
                s.dump()
            p.time = float(p.current_dumps) * p.final_sim_time / float(p.num_data_dumps)
            if p.expanding:
                p.tau = get_tau(p.time, p.cosmo_params)
        p.n_steps += 1

    def not_finished(self) -> bool:
        return self.head.not_finished()
