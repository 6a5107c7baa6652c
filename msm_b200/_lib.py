"""ctypes binding of libmsm_b200.so (the C ABI declared in include/msm_b200.h).

There is no fallback of any kind: if the shared library is missing, importing this module raises, and if no
sm_100 GPU is usable every compute entry point returns MSM_E_CUDA which is raised as `MsmError`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSM_B200_LIB", os.path.join(_HERE, "libmsm_b200.so"))   # override: A/B builds only

MSM_OK = 0
MSM_E_ARG, MSM_E_CUDA, MSM_E_NCCL, MSM_E_ALIASING, MSM_E_NAN, MSM_E_STATE, MSM_E_NOMEM, MSM_E_IO = \
    -1, -2, -3, -4, -5, -6, -7, -8
COUPLING_INDEPENDENT, COUPLING_SUMMED = 0, 1
SCHEME_NONE, SCHEME_POISSON, SCHEME_WIGNER, SCHEME_HUSIMI = 0, 1, 2, 3
SCHEMES = {"Poisson": SCHEME_POISSON, "Wigner": SCHEME_WIGNER, "Husimi": SCHEME_HUSIMI}
SEED_NONE = 2 ** 64 - 1


class MsmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"msm_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class MsmConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("dims", C.c_int32), ("size", C.c_int32), ("n_streams", C.c_int32),
                ("coupling", C.c_int32), ("device", C.c_int32), ("chunk_streams", C.c_int32), ("rank", C.c_int32),
                ("nranks", C.c_int32), ("n_streams_global", C.c_int32), ("dx", C.c_double),
                ("density_prefactor", C.c_double), ("poisson_coeff", C.c_double), ("k2_cutoff", C.c_double),
                ("nccl_unique_id", C.c_void_p)]


class MsmSimParams(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("dims", C.c_int32), ("size", C.c_int32), ("n_streams", C.c_int32),
                ("expanding", C.c_int32), ("coupling", C.c_int32), ("device", C.c_int32),
                ("chunk_streams", C.c_int32), ("num_data_dumps", C.c_uint32), ("has_max_dloga", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("n_streams_global", C.c_int32), ("reserved", C.c_int32),
                ("axis_length", C.c_double), ("time", C.c_double), ("final_sim_time", C.c_double),
                ("cfl", C.c_double), ("total_mass", C.c_double), ("particle_mass", C.c_double),
                ("hbar_", C.c_double), ("k2_cutoff", C.c_double), ("alias_threshold", C.c_double),
                ("omega_matter_now", C.c_double), ("omega_radiation_now", C.c_double), ("h", C.c_double),
                ("z0", C.c_double), ("max_dloga", C.c_double), ("nccl_unique_id", C.c_void_p)]


class MsmStreamState(C.Structure):
    _fields_ = [("time", C.c_double), ("tau", C.c_double), ("dt", C.c_double), ("potential_max", C.c_double),
                ("alias_mass", C.c_double), ("scale_factor", C.c_double), ("n_steps", C.c_uint64),
                ("current_dumps", C.c_uint32), ("dumped", C.c_int32), ("finished", C.c_int32),
                ("aliased", C.c_int32)]


class MsmDerived(C.Structure):
    _fields_ = [("dx", C.c_double), ("dk", C.c_double), ("k2_max", C.c_double), ("n_tot", C.c_double),
                ("comoving_boxsize", C.c_double), ("tau0", C.c_double), ("final_sim_tau", C.c_double),
                ("density_prefactor", C.c_double), ("poisson_coeff", C.c_double)]


class MsmProfileRecord(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_uint64), ("ms_total", C.c_double),
                ("algorithmic_bytes", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against include/msm_b200.h
PROTOTYPES = {
    "msm_version": (C.c_char_p, []),
    "msm_strerror": (C.c_char_p, [C.c_int]),
    "msm_nccl_unique_id": (C.c_int, [_vp]),
    "msm_create": (C.c_int, [C.POINTER(MsmConfig), C.POINTER(_vp)]),
    "msm_destroy": (None, [_vp]),
    "msm_last_error": (C.c_char_p, [_vp]),
    "msm_device_bytes": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "msm_set_psi": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_set_psi_planes": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_get_psi": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_get_psi_interleaved": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_get_psi_many": (C.c_int, [_vp, C.c_int32, _ip, C.POINTER(_dp), C.POINTER(_dp)]),
    "msm_upload_begin": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_download_begin": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_transfers_wait": (C.c_int, [_vp]),
    "msm_chunk_streams": (C.c_int, [_vp, _ip]),
    "msm_download_ticket": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "msm_download_wait": (C.c_int, [_vp, C.c_uint64]),
    "msm_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "msm_host_free": (C.c_int, [_vp, _vp]),
    "msm_get_psik_interleaved": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_potential_max": (C.c_int, [_vp, _ip, _dp]),
    "msm_get_potential": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_step": (C.c_int, [_vp, _ip, _dp, _dp, _dp]),
    "msm_read_alias": (C.c_int, [_vp, _dp]),
    "msm_allreduce_max": (C.c_int, [_vp, _dp]),
    "msm_synchronize": (C.c_int, [_vp]),
    "msm_fft": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _dp]),
    "msm_spec_grid": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp]),
    "msm_ic_cold_gauss": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_ic_spherical_tophat": (C.c_int, [_vp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]),
    "msm_ic_cold_gauss_kspace": (C.c_int, [_vp, C.c_int32, _dp, _dp, C.c_uint64]),
    "msm_ic_copy": (C.c_int, [_vp, C.c_int32, C.c_int32]),
    "msm_ic_store": (C.c_int, [_vp, C.c_int32]),
    "msm_ic_load": (C.c_int, [_vp, C.c_int32]),
    "msm_sample_perturbation": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_uint64, C.c_double]),
    "msm_ensemble_accumulate": (C.c_int, [_vp, _ip]),
    "msm_ensemble_allreduce": (C.c_int, [_vp]),
    "msm_ensemble_get": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_profile_enable": (C.c_int, [_vp, C.c_int32]),
    "msm_profile_read": (C.c_int, [_vp, C.POINTER(MsmProfileRecord), C.c_int32, _ip]),
    "msm_timer_start": (C.c_int, [_vp]),
    "msm_timer_stop": (C.c_int, [_vp, _dp]),
    "msm_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "msm_sim_create": (C.c_int, [C.POINTER(MsmSimParams), C.POINTER(_vp)]),
    "msm_sim_destroy": (None, [_vp]),
    "msm_sim_last_error": (C.c_char_p, [_vp]),
    "msm_sim_ctx": (_vp, [_vp]),
    "msm_sim_derived": (C.c_int, [_vp, C.POINTER(MsmDerived)]),
    "msm_sim_set_psi": (C.c_int, [_vp, C.c_int32, _dp]),
    "msm_sim_update": (C.c_int, [_vp]),
    "msm_sim_update_streams": (C.c_int, [_vp, _ip]),
    "msm_sim_run_streams": (C.c_int, [_vp, C.c_int32, _ip, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.c_uint64]),
    "msm_sim_run_streams_seeded": (C.c_int, [_vp, C.c_int32, _ip, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(_dp),
                                           C.POINTER(_dp), C.c_uint64]),
    "msm_run_groups": (C.c_int, [C.c_int32, C.c_int32, _ip, C.c_int32]),
    "msm_sim_not_finished": (C.c_int, [_vp]),
    "msm_sim_state": (C.c_int, [_vp, C.c_int32, C.POINTER(MsmStreamState)]),
    "msm_sim_get_psi": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "msm_sim_dump": (C.c_int, [_vp, C.c_int32, C.c_char_p, C.c_char_p, C.c_uint32]),
    "msm_sim_dump_potential": (C.c_int, [_vp, C.c_int32, C.c_char_p, C.c_char_p, C.c_uint32]),
    "msm_sim_wait_io": (C.c_int, [_vp]),
    "msm_sim_reserve_dump_buffers": (C.c_int, [_vp, C.c_int32]),
    "msm_get_tau": (C.c_double, [C.c_double] * 6 + [C.c_int32]),
    "msm_supercomoving_boxsize": (C.c_double, [C.c_double] * 5),
    "msm_scale_factor_after": (C.c_double, [C.c_double] * 6),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C msm_b200/csrc`).  msm_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(code: int, handle=None, sim: bool = False) -> None:
    if code == MSM_OK:
        return
    if sim:
        msg = lib.msm_sim_last_error(handle)
    else:
        msg = lib.msm_last_error(handle)
    text = (msg or b"").decode("utf-8", "replace") or lib.msm_strerror(code).decode()
    raise MsmError(code, text)
