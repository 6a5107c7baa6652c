"""Python host over the C ABI: thin object wrappers whose method names follow the reference.

`Context`            grid level (msm_*): the device arrays of `SimulationGrid` (simulation_object.rs:42-64)
`SimulationObject`   host-logic level (msm_sim_*): mirror of the reference's `SimulationObject`
                     (simulation_object.rs:145-184) batched over the streams of one TOML
`forward` / `inverse` / `spec_grid`   the FFT layer of utils/fft.rs on host arrays

NumPy is used only to hold host buffers; all arithmetic happens in libmsm_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (COUPLING_INDEPENDENT, COUPLING_SUMMED, MsmConfig, MsmDerived, MsmError, MsmProfileRecord,
                   MsmSimParams, MsmStreamState, check, lib)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _f64(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _as_interleaved(psi: np.ndarray) -> np.ndarray:
    """complex128 C-order array -> flat float64 view (re, im, re, im, ...)."""
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    return psi.reshape(-1).view(np.float64)


# --------------------------------------------------------------------------------------------------------------
# FFT layer (utils/fft.rs)
# --------------------------------------------------------------------------------------------------------------
def _fft(a: np.ndarray, dims: int, inverse_: bool, device: int) -> np.ndarray:
    a = np.array(a, dtype=np.complex128, order="C", copy=True)
    size = a.shape[-1]
    if any(s != size for s in a.shape[-dims:]):
        raise ValueError("only uniform grids are supported")
    batch = int(np.prod(a.shape[:-dims])) if a.ndim > dims else 1
    flat = a.reshape(-1).view(np.float64)
    check(lib.msm_fft(device, dims, size, 1 if inverse_ else 0, batch, _f64(flat)))
    return a


def forward(a: np.ndarray, dims: Optional[int] = None, device: int = 0) -> np.ndarray:
    """utils/fft.rs:6-31 `forward`: unitary d-dim DFT over the last `dims` axes (leading axes = batch)."""
    return _fft(a, a.ndim if dims is None else dims, False, device)


def inverse(a: np.ndarray, dims: Optional[int] = None, device: int = 0) -> np.ndarray:
    """utils/fft.rs:33-58 `inverse`."""
    return _fft(a, a.ndim if dims is None else dims, True, device)


def spec_grid(dx: float, dims: int, size: int, device: int = 0) -> np.ndarray:
    """utils/fft.rs:123-161 `spec_grid`, evaluated by the same device expression the kernels use."""
    out = np.empty((size,) * dims, dtype=np.float64)
    check(lib.msm_spec_grid(device, dims, size, float(dx), _f64(out.reshape(-1))))
    return out


# --------------------------------------------------------------------------------------------------------------
# grid level
# --------------------------------------------------------------------------------------------------------------
class Context:
    """Device state of a batch of streams (msm_create ... msm_destroy)."""

    def __init__(self, dims: int, size: int, n_streams: int, dx: float, density_prefactor: float,
                 poisson_coeff: float, k2_cutoff: float = 0.95, coupling: int = COUPLING_INDEPENDENT,
                 device: int = 0, chunk_streams: int = 0, rank: int = 0, nranks: int = 1,
                 n_streams_global: int = 0, nccl_unique_id: Optional[bytes] = None):
        self._uid = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id else None
        cfg = MsmConfig(C.sizeof(MsmConfig), dims, size, n_streams, coupling, device, chunk_streams, rank, nranks,
                        n_streams_global, dx, density_prefactor, poisson_coeff, k2_cutoff,
                        C.cast(self._uid, C.c_void_p) if self._uid else None)
        self.handle = C.c_void_p()
        check(lib.msm_create(C.byref(cfg), C.byref(self.handle)))
        self.dims, self.size, self.n_streams = dims, size, n_streams
        self.shape = (size,) * dims
        self.cells = size ** dims
        self._owned = True

    @classmethod
    def _borrow(cls, handle, dims, size, n_streams):
        self = cls.__new__(cls)
        self.handle = C.c_void_p(handle)
        self.dims, self.size, self.n_streams = dims, size, n_streams
        self.shape = (size,) * dims
        self.cells = size ** dims
        self._owned = False
        return self

    def close(self):
        if getattr(self, "_owned", False) and self.handle:
            lib.msm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, code):
        check(code, self.handle)

    def device_bytes(self) -> int:
        v = C.c_uint64()
        self._chk(lib.msm_device_bytes(self.handle, C.byref(v)))
        return v.value

    def set_psi(self, stream: int, psi: np.ndarray) -> None:
        flat = _as_interleaved(psi)
        assert flat.size == 2 * self.cells
        self._chk(lib.msm_set_psi(self.handle, stream, _f64(flat)))

    def set_psi_planes(self, stream: int, re: np.ndarray, im: np.ndarray) -> None:
        re = np.ascontiguousarray(re, dtype=np.float64).reshape(-1)
        im = np.ascontiguousarray(im, dtype=np.float64).reshape(-1)
        assert re.size == self.cells and im.size == self.cells
        self._chk(lib.msm_set_psi_planes(self.handle, stream, _f64(re), _f64(im)))

    def get_psi(self, stream: int) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.complex128)
        self._chk(lib.msm_get_psi_interleaved(self.handle, stream, _f64(out.reshape(-1).view(np.float64))))
        return out

    def get_psi_planes(self, stream: int):
        re = np.empty(self.shape, dtype=np.float64)
        im = np.empty(self.shape, dtype=np.float64)
        self._chk(lib.msm_get_psi(self.handle, stream, _f64(re.reshape(-1)), _f64(im.reshape(-1))))
        return re, im

    def get_psi_many(self, streams: Sequence[int], re: Sequence[np.ndarray], im: Sequence[np.ndarray]) -> None:
        """Pipelined dump (row f-2): planes of stream i cross PCIe while stream i+1 is transformed.  `re[i]`, `im[i]`
        are float64 arrays of n^dims elements (pinned memory gives real overlap); the same array may be passed twice."""
        n = len(streams)
        ids = np.ascontiguousarray(streams, dtype=np.int32)
        rp = (_dp * n)(*[_f64(a.reshape(-1)) for a in re])
        ip = (_dp * n)(*[_f64(a.reshape(-1)) for a in im])
        self._chk(lib.msm_get_psi_many(self.handle, n, ids.ctypes.data_as(_ip), rp, ip))

    # asynchronous transfers: the arrays must stay alive (and unmodified) until transfers_wait() returns
    def upload_begin(self, stream: int, psi_flat: np.ndarray) -> None:
        assert psi_flat.size == 2 * self.cells
        self._chk(lib.msm_upload_begin(self.handle, stream, _f64(psi_flat)))

    def download_begin(self, stream: int, re: Optional[np.ndarray], im: Optional[np.ndarray]) -> None:
        self._chk(lib.msm_download_begin(self.handle, stream, _f64(re.reshape(-1)) if re is not None else None,
                                         _f64(im.reshape(-1)) if im is not None else None))

    def transfers_wait(self) -> None:
        self._chk(lib.msm_transfers_wait(self.handle))

    def chunk_streams(self) -> int:
        v = C.c_int32()
        self._chk(lib.msm_chunk_streams(self.handle, C.byref(v)))
        return v.value

    def get_psik(self, stream: int) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.complex128)
        self._chk(lib.msm_get_psik_interleaved(self.handle, stream, _f64(out.reshape(-1).view(np.float64))))
        return out

    def _active(self, active):
        if active is None:
            return None, None
        arr = np.ascontiguousarray(active, dtype=np.int32)
        assert arr.size == self.n_streams
        return arr, arr.ctypes.data_as(_ip)

    def potential_max(self, active=None) -> np.ndarray:
        out = np.zeros(self.n_streams, dtype=np.float64)
        keep, ptr = self._active(active)
        self._chk(lib.msm_potential_max(self.handle, ptr, _f64(out)))
        return out

    def get_potential(self, stream: int) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        self._chk(lib.msm_get_potential(self.handle, stream, _f64(out.reshape(-1))))
        return out

    def step(self, drift_coeff: Sequence[float], kick_coeff: Sequence[float], active=None,
             blocking: bool = True) -> Optional[np.ndarray]:
        d = np.ascontiguousarray(drift_coeff, dtype=np.float64)
        k = np.ascontiguousarray(kick_coeff, dtype=np.float64)
        assert d.size == self.n_streams and k.size == self.n_streams
        keep, ptr = self._active(active)
        if blocking:
            alias = np.zeros(self.n_streams, dtype=np.float64)
            self._chk(lib.msm_step(self.handle, ptr, _f64(d), _f64(k), _f64(alias)))
            return alias
        self._chk(lib.msm_step(self.handle, ptr, _f64(d), _f64(k), None))
        return None

    def read_alias(self) -> np.ndarray:
        alias = np.zeros(self.n_streams, dtype=np.float64)
        self._chk(lib.msm_read_alias(self.handle, _f64(alias)))
        return alias

    def synchronize(self) -> None:
        self._chk(lib.msm_synchronize(self.handle))

    # on-device initial conditions
    def ic_cold_gauss(self, stream: int, mean: Sequence[float], std: Sequence[float]) -> None:
        m = np.ascontiguousarray(mean, dtype=np.float64)
        s = np.ascontiguousarray(std, dtype=np.float64)
        assert m.size == self.dims and s.size == self.dims
        self._chk(lib.msm_ic_cold_gauss(self.handle, stream, _f64(m), _f64(s)))

    def ic_spherical_tophat(self, stream: int, axis_length: float, radius: float, delta: float, slope: float) -> None:
        self._chk(lib.msm_ic_spherical_tophat(self.handle, stream, axis_length, radius, delta, slope))

    def ic_cold_gauss_kspace(self, stream: int, mean: Sequence[float], std: Sequence[float], phase_seed: int = 0) -> None:
        m = np.ascontiguousarray(mean, dtype=np.float64)
        s = np.ascontiguousarray(std, dtype=np.float64)
        assert m.size == self.dims and s.size == self.dims
        self._chk(lib.msm_ic_cold_gauss_kspace(self.handle, stream, _f64(m), _f64(s), int(phase_seed)))

    def ic_copy(self, dst: int, src: int) -> None:
        self._chk(lib.msm_ic_copy(self.handle, dst, src))

    def ic_store(self, stream: int) -> None:
        """keep the stream's wavefunction aside (the un-sampled IC every stream starts from)"""
        self._chk(lib.msm_ic_store(self.handle, stream))

    def ic_load(self, stream: int) -> None:
        self._chk(lib.msm_ic_load(self.handle, stream))

    def sample_perturbation(self, stream: int, scheme: str, seed: int, n_tot: float) -> None:
        self._chk(lib.msm_sample_perturbation(self.handle, stream, _lib.SCHEMES[scheme], int(seed), float(n_tot)))

    # ensemble statistics (the synthesizer's stream reductions, on the device)
    def ensemble_sums(self, active=None, allreduce: bool = False) -> Dict[str, np.ndarray]:
        """SUMS over the selected streams of psi, |psi|^2, psi_k (un-normalised DFT), |psi_k|^2
        (synthesizer/src/main.rs:63-93); divide by the global stream count for the synthesizer's means.
        allreduce: also sum over the ranks of the communicator (ncclAllReduce of the four grids, in place)."""
        keep, ptr = self._active(active)
        self._chk(lib.msm_ensemble_accumulate(self.handle, ptr))
        if allreduce:
            self._chk(lib.msm_ensemble_allreduce(self.handle))
        out = {}
        for field, name in enumerate(("psi", "psi2", "psik", "psik2")):
            re = np.empty(self.shape, dtype=np.float64)
            im = np.empty(self.shape, dtype=np.float64)
            self._chk(lib.msm_ensemble_get(self.handle, field, _f64(re.reshape(-1)), _f64(im.reshape(-1))))
            out[name] = re + 1j * im
        return out

    # profiling
    def profile_enable(self, on: bool = True) -> None:
        self._chk(lib.msm_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self) -> List[Dict]:
        n = C.c_int32()
        self._chk(lib.msm_profile_read(self.handle, None, 0, C.byref(n)))
        recs = (MsmProfileRecord * max(1, n.value))()
        self._chk(lib.msm_profile_read(self.handle, recs, n.value, C.byref(n)))
        return [dict(name=recs[i].name.decode(), launches=int(recs[i].launches), ms_total=float(recs[i].ms_total),
                     algorithmic_bytes=float(recs[i].algorithmic_bytes)) for i in range(n.value)]

    def timer_start(self) -> None:
        self._chk(lib.msm_timer_start(self.handle))

    def timer_stop(self) -> float:
        v = C.c_double()
        self._chk(lib.msm_timer_stop(self.handle, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        v = C.c_uint64()
        self._chk(lib.msm_launch_count(self.handle, C.byref(v)))
        return v.value


# --------------------------------------------------------------------------------------------------------------
# host-logic level
# --------------------------------------------------------------------------------------------------------------
@dataclass
class CosmologyParameters:
    """common/src/parameters.rs:71-86."""
    omega_matter_now: float
    omega_radiation_now: float
    h: float
    z0: float
    max_dloga: Optional[float] = None


@dataclass
class SimulationParameters:
    """The resolved scalars of simulation_object.rs:67-140 shared by all streams of a run."""
    axis_length: float
    final_sim_time: float
    cfl: float
    num_data_dumps: int
    total_mass: float
    particle_mass: float
    hbar_: float
    k2_cutoff: float
    alias_threshold: float
    dims: int
    size: int
    time: float = 0.0
    cosmology: Optional[CosmologyParameters] = None     # None <=> static box (cargo feature `expanding` off)


class FourierAliasing(MsmError):
    """utils/error.rs `RuntimeError::FourierAliasing`; the reference panics (simulation_object.rs:607-617)."""


class SimulationObject:
    """Mirror of the reference's `SimulationObject` for `n_streams` streams advanced together.

    update() / not_finished() / get_timestep-derived state / dump() keep the reference's meaning
    (simulation_object.rs:475, :669, :1226, :1113); streams are independent unless coupling == COUPLING_SUMMED."""

    def __init__(self, parameters: SimulationParameters, n_streams: int = 1, coupling: int = COUPLING_INDEPENDENT,
                 device: int = 0, chunk_streams: int = 0, rank: int = 0, nranks: int = 1, n_streams_global: int = 0,
                 nccl_unique_id: Optional[bytes] = None):
        p = parameters
        c = p.cosmology
        self._uid = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id else None
        sp = MsmSimParams()
        sp.struct_size = C.sizeof(MsmSimParams)
        sp.dims, sp.size, sp.n_streams = p.dims, p.size, n_streams
        sp.expanding = 1 if c is not None else 0
        sp.coupling, sp.device, sp.chunk_streams = coupling, device, chunk_streams
        sp.num_data_dumps = p.num_data_dumps
        sp.rank, sp.nranks, sp.n_streams_global = rank, nranks, n_streams_global
        sp.axis_length, sp.time, sp.final_sim_time, sp.cfl = p.axis_length, p.time, p.final_sim_time, p.cfl
        sp.total_mass, sp.particle_mass, sp.hbar_ = p.total_mass, p.particle_mass, p.hbar_
        sp.k2_cutoff, sp.alias_threshold = p.k2_cutoff, p.alias_threshold
        if c is not None:
            sp.omega_matter_now, sp.omega_radiation_now, sp.h, sp.z0 = c.omega_matter_now, c.omega_radiation_now, c.h, c.z0
            sp.has_max_dloga = 1 if c.max_dloga is not None else 0
            sp.max_dloga = c.max_dloga if c.max_dloga is not None else 0.0
        sp.nccl_unique_id = C.cast(self._uid, C.c_void_p) if self._uid else None
        self.handle = C.c_void_p()
        check(lib.msm_sim_create(C.byref(sp), C.byref(self.handle)), None, sim=True)
        self.parameters = parameters
        self.n_streams = n_streams
        self.shape = (p.size,) * p.dims
        self.grid = Context._borrow(lib.msm_sim_ctx(self.handle), p.dims, p.size, n_streams)
        d = MsmDerived()
        check(lib.msm_sim_derived(self.handle, C.byref(d)), self.handle, sim=True)
        self.derived = d

    def close(self):
        if self.handle:
            lib.msm_sim_destroy(self.handle)
        self.handle = None
        self.grid = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_psi(self, stream: int, psi: np.ndarray) -> None:
        flat = _as_interleaved(psi)
        check(lib.msm_sim_set_psi(self.handle, stream, _f64(flat)), self.handle, sim=True)

    def update(self, raise_on_alias: bool = True) -> None:
        """One `update()` for every unfinished stream."""
        code = lib.msm_sim_update(self.handle)
        if code == _lib.MSM_E_ALIASING:
            if raise_on_alias:
                raise FourierAliasing(code, lib.msm_sim_last_error(self.handle).decode())
            return
        check(code, self.handle, sim=True)

    def update_streams(self, subset: Sequence[int], raise_on_alias: bool = True) -> None:
        """One `update()` for the unfinished streams with subset[s] != 0."""
        arr = np.ascontiguousarray(subset, dtype=np.int32)
        assert arr.size == self.n_streams
        code = lib.msm_sim_update_streams(self.handle, arr.ctypes.data_as(_ip))
        if code == _lib.MSM_E_ALIASING:
            if raise_on_alias:
                raise FourierAliasing(code, lib.msm_sim_last_error(self.handle).decode())
            return
        check(code, self.handle, sim=True)

    def run_streams(self, streams: Sequence[int], psi_in: Optional[Sequence[Optional[np.ndarray]]] = None,
                    re_out: Optional[Sequence[Optional[np.ndarray]]] = None,
                    im_out: Optional[Sequence[Optional[np.ndarray]]] = None, max_updates: int = 0,
                    raise_on_alias: bool = True) -> None:
        """The reference's outer loop (simulator/src/main.rs:43-85) for the listed streams: upload psi_in[i] (flat
        float64 views of complex128 grids), `while not_finished(): update()` (at most max_updates per stream when > 0),
        final psi into re_out[i] / im_out[i].  Transfers of neighbouring stream groups overlap the step kernels."""
        n = len(streams)
        ids = np.ascontiguousarray(streams, dtype=np.int32)

        def ptrs(arrs):
            if arrs is None:
                return None
            assert len(arrs) == n
            return (_dp * n)(*[_f64(a.reshape(-1)) if a is not None else None for a in arrs])
        for a in (psi_in or []):
            assert a is None or a.size == 2 * self.grid.cells
        code = lib.msm_sim_run_streams(self.handle, n, ids.ctypes.data_as(_ip), ptrs(psi_in), ptrs(re_out), ptrs(im_out),
                                       int(max_updates))
        if code == _lib.MSM_E_ALIASING:
            if raise_on_alias:
                raise FourierAliasing(code, lib.msm_sim_last_error(self.handle).decode())
            return
        check(code, self.handle, sim=True)

    def run_streams_seeded(self, streams: Sequence[int], scheme: Optional[str], seeds: Sequence[Optional[int]],
                           re_out: Optional[Sequence[Optional[np.ndarray]]] = None,
                           im_out: Optional[Sequence[Optional[np.ndarray]]] = None, max_updates: int = 0,
                           raise_on_alias: bool = True) -> None:
        """run_streams with the initial conditions built on the device: the wavefunction saved by `grid.ic_store`
        + `sample_quantum_perturbation` with seeds[i] (None = un-sampled).  Only scalars cross PCIe on the way in."""
        n = len(streams)
        ids = np.ascontiguousarray(streams, dtype=np.int32)
        sd = np.array([_lib.SEED_NONE if s is None else int(s) for s in seeds], dtype=np.uint64)
        assert sd.size == n

        def ptrs(arrs):
            if arrs is None:
                return None
            assert len(arrs) == n
            return (_dp * n)(*[_f64(a.reshape(-1)) if a is not None else None for a in arrs])
        code = lib.msm_sim_run_streams_seeded(self.handle, n, ids.ctypes.data_as(_ip), _lib.SCHEMES.get(scheme, 0),
                                              sd.ctypes.data_as(C.POINTER(C.c_uint64)), ptrs(re_out), ptrs(im_out),
                                              int(max_updates))
        if code == _lib.MSM_E_ALIASING:
            if raise_on_alias:
                raise FourierAliasing(code, lib.msm_sim_last_error(self.handle).decode())
            return
        check(code, self.handle, sim=True)

    def not_finished(self) -> bool:
        return bool(lib.msm_sim_not_finished(self.handle))

    def state(self, stream: int) -> MsmStreamState:
        st = MsmStreamState()
        check(lib.msm_sim_state(self.handle, stream, C.byref(st)), self.handle, sim=True)
        return st

    def get_psi(self, stream: int) -> np.ndarray:
        re = np.empty(self.shape, dtype=np.float64)
        im = np.empty(self.shape, dtype=np.float64)
        check(lib.msm_sim_get_psi(self.handle, stream, _f64(re.reshape(-1)), _f64(im.reshape(-1))), self.handle, sim=True)
        return re + 1j * im

    def dump(self, stream: int, root_dir: str, sim_name: str, dump_index: int) -> None:
        """`complex_array_to_disk` layout: <root>/<sim_name>/psi_%05d_real|_imag (utils/io.rs:34-88)."""
        check(lib.msm_sim_dump(self.handle, stream, root_dir.encode(), sim_name.encode(), dump_index), self.handle, sim=True)

    def dump_potential(self, stream: int, root_dir: str, sim_name: str, dump_index: int) -> None:
        """`output_potential` (simulation_object.rs:1167-1180): <root>/<sim_name>/potential_%05d_real|_imag (imag = 0)."""
        check(lib.msm_sim_dump_potential(self.handle, stream, root_dir.encode(), sim_name.encode(), dump_index),
              self.handle, sim=True)

    def reserve_dump_buffers(self, n: int) -> None:
        """pin n staging buffers of the dump pipeline up front (otherwise the pool grows at the first dumps)"""
        check(lib.msm_sim_reserve_dump_buffers(self.handle, int(n)), self.handle, sim=True)

    def wait_io(self) -> None:
        check(lib.msm_sim_wait_io(self.handle), self.handle, sim=True)


def get_tau(target_time: float, c: CosmologyParameters) -> float:
    """simulation_object.rs:1408-1453 `get_tau` (host scalar)."""
    return float(lib.msm_get_tau(target_time, c.omega_matter_now, c.omega_radiation_now, c.h, c.z0,
                                 c.max_dloga if c.max_dloga is not None else 0.0, 1 if c.max_dloga is not None else 0))


def get_supercomoving_boxsize(hbar_: float, c: CosmologyParameters, axis_length: float) -> float:
    """common/src/parameters.rs:205-220."""
    return float(lib.msm_supercomoving_boxsize(hbar_, c.omega_matter_now, c.h, c.z0, axis_length))


def scale_factor_after(t: float, c: CosmologyParameters) -> float:
    """expanding.rs:99-105 `ScaleFactorSolver::step(t)` from t0 = 0."""
    return float(lib.msm_scale_factor_after(t, c.omega_matter_now, c.omega_radiation_now, c.h, c.z0,
                                            c.max_dloga if c.max_dloga is not None else 1e-3))
