"""ctypes binding of libbiolith_b200.so (the C ABI in include/biolith_b200.h).

There is deliberately NO fallback: if the shared object is missing or a call fails, this raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbiolith_b200.so")

BL_ABI_VERSION = 1
BL_MODEL = {"occu": 0, "occu_rn": 1, "occu_cop": 2, "nmixture": 3, "occu_cs": 4}
BL_F32, BL_F64 = 0, 1
BL_FLAG_FP_CONSTANT, BL_FLAG_FP_UNOCCUPIED, BL_FLAG_PRIOR, BL_FLAG_STRICT_MATH = 1, 2, 4, 8
BL_FLAG_SITE_RE, BL_FLAG_OBS_RE = 16, 32


class BiolithB200Error(RuntimeError):
    def __init__(self, status, what, detail):
        super().__init__(f"{what}: {detail} [status {status}]")
        self.status = status


class bl_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("dtype", C.c_int32), ("data_dtype", C.c_int32),
        ("flags", C.c_uint32), ("device", C.c_int32), ("n_sites", C.c_int64), ("n_periods", C.c_int32),
        ("n_replicates", C.c_int32), ("n_site_covs", C.c_int32), ("n_obs_covs", C.c_int32),
        ("n_species", C.c_int32), ("max_abundance", C.c_int32), ("max_chains", C.c_int32),
        ("reserved0", C.c_int32),
        ("prior_beta_loc", C.c_double), ("prior_beta_scale", C.c_double),
        ("prior_alpha_loc", C.c_double), ("prior_alpha_scale", C.c_double),
        ("prior_fp_a", C.c_double), ("prior_fp_b", C.c_double), ("prior_fp_rate", C.c_double),
    ]


class bl_info(C.Structure):
    _fields_ = [
        ("theta_dim", C.c_int32), ("n_extras", C.c_int32), ("n_units", C.c_int64),
        ("packed_bytes", C.c_int64), ("algorithmic_bytes", C.c_int64), ("n_masked", C.c_int64),
        ("fields_per_unit", C.c_int32), ("kernel_variant", C.c_int32),
    ]


class bl_xla_opaque(C.Structure):
    _fields_ = [("dataset", C.c_uint64), ("n_chains", C.c_int32), ("reserved", C.c_int32)]


class bl_nuts_config(C.Structure):
    _fields_ = [
        ("n_chains", C.c_int32), ("num_warmup", C.c_int32), ("num_samples", C.c_int32),
        ("max_tree_depth", C.c_int32), ("adapt_step_size", C.c_int32), ("adapt_mass_matrix", C.c_int32),
        ("find_heuristic_step_size", C.c_int32), ("reserved0", C.c_int32), ("seed", C.c_uint64), ("target_accept_prob", C.c_double), ("init_step_size", C.c_double),
        ("max_delta_energy", C.c_double),
    ]


_P = C.c_void_p
# name -> (restype, argtypes); every symbol include/biolith_b200.h declares
SIGNATURES = {
    "bl_version": (C.c_int, []),
    "bl_strerror": (C.c_char_p, [C.c_int]),
    "bl_last_error": (C.c_char_p, []),
    "bl_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bl_launch_count": (C.c_int64, []),
    "bl_dataset_create": (C.c_int, [C.POINTER(bl_desc), _P, _P, _P, _P, C.POINTER(_P)]),
    "bl_dataset_destroy": (C.c_int, [_P]),
    "bl_dataset_info": (C.c_int, [_P, C.POINTER(bl_info)]),
    "bl_plan_kernel": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]),
    "bl_dataset_export_mask": (C.c_int, [_P, _P]),
    "bl_eval": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "bl_eval_host": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "bl_eval_timed": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_int32, C.POINTER(C.c_float)]),
    "bl_site_summary": (C.c_int, [_P, _P, C.c_int32, _P]),
    "bl_obs_loglik": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "bl_nuts_create": (C.c_int, [_P, C.POINTER(bl_nuts_config), _P, C.POINTER(_P)]),
    "bl_nuts_run": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "bl_nuts_get": (C.c_int, [_P] * 11),
    "bl_nuts_rows_evaluated": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "bl_nuts_destroy": (C.c_int, [_P]),
    "bl_comm_unique_id": (C.c_int, [_P, C.c_size_t]),
    "bl_dataset_attach_nccl": (C.c_int, [_P, _P, C.c_size_t, C.c_int32, C.c_int32]),
    "bl_dataset_p2p_export": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_size_t]),
    "bl_dataset_p2p_attach": (C.c_int, [_P, _P, C.c_size_t]),
    "bl_dataset_comm_error": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "bl_dataset_detach_comm": (C.c_int, [_P]),
    "bl_xla_eval": (None, [_P, C.POINTER(_P), C.c_char_p, C.c_size_t, _P]),
    "bl_device_malloc": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(_P)]),
    "bl_device_free": (C.c_int, [_P]),
    "bl_memcpy_h2d": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "bl_memcpy_d2h": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "bl_host_malloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "bl_host_free_pinned": (C.c_int, [_P]),
    "bl_stream_create": (C.c_int, [C.c_int32, C.POINTER(_P)]),
    "bl_stream_destroy": (C.c_int, [_P]),
    "bl_stream_sync": (C.c_int, [_P]),
    "bl_flush_l2": (C.c_int, [C.c_int32, _P]),
    "bl_pipe_peak": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bl_mufu_error": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int64, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the shared object (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BiolithB200Error(
                -3, "libbiolith_b200.so not found",
                f"{LIB_PATH} is missing; run `python -m biolith_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.bl_version() != BL_ABI_VERSION:
            raise BiolithB200Error(-1, "ABI mismatch", f"library {lib.bl_version()} != binding {BL_ABI_VERSION}")
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != 0:
        lib = load()
        detail = lib.bl_last_error().decode() or lib.bl_strerror(status).decode()
        raise BiolithB200Error(status, what, detail)
