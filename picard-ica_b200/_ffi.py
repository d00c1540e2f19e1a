"""ctypes binding of libpicard_b200.so (include/picard_b200.h).  Loading fails LOUDLY when the CUDA
library has not been built: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PICARD_B200_LIB") or os.path.join(_HERE, "libpicard_b200.so")  # env override: profiling builds only

dp = C.POINTER(C.c_double)

# every symbol include/picard_b200.h declares (tests check the library exports all of them)
EXPORTED = [
    "picard_abi_version", "picard_device_count", "picard_status_string", "picard_release_cache", "picard_config_default", "picard_config_validate",
    "picard_fit", "picard_fit_device", "picard_transform", "picard_result_free", "picard_core_create", "picard_core_run",
    "picard_core_reset", "picard_core_state", "picard_core_stats", "picard_core_destroy", "picard_eval_moments", "picard_eval_moments_ex",
    "picard_eval_moments_device", "picard_eval_moments_device_ex",
    "picard_eval_point", "picard_matrix_exp", "picard_sln_det", "picard_sym_decorrelation", "picard_compute_direction",
    "picard_center_whiten", "picard_center_whiten_device", "picard_jade", "picard_jade_cumulants", "picard_synth_sources", "picard_fp64_peak_probe", "picard_apply_device", "picard_comm_unique_id",
    "picard_comm_create", "picard_comm_rank", "picard_comm_size", "picard_comm_destroy",
]


class Config(C.Structure):  # picard_config_t
    _fields_ = [
        ("density_kind", C.c_int32), ("alpha", C.c_double), ("n_components", C.c_int64),
        ("ortho", C.c_int32), ("extended", C.c_int32), ("whiten", C.c_int32), ("centering", C.c_int32),
        ("max_iter", C.c_int64), ("tol", C.c_double), ("m", C.c_int64), ("ls_tries", C.c_int64), ("lambda_min", C.c_double),
        ("w_init", dp), ("w_init_rows", C.c_int64), ("w_init_cols", C.c_int64),
        ("fastica_it", C.c_int64), ("jade_it", C.c_int64), ("has_seed", C.c_int32), ("seed", C.c_uint64), ("verbose", C.c_int32),
        ("device", C.c_int32), ("comm", C.c_void_p), ("flags", C.c_uint32),
    ]


class Stats(C.Structure):  # picard_stats_t
    _fields_ = [
        ("core_ms", C.c_double), ("preprocess_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("fused_passes", C.c_int64), ("grad_passes", C.c_int64), ("loss_passes", C.c_int64),
        ("ls_tries", C.c_int64), ("fallbacks", C.c_int64), ("sign_changes", C.c_int64), ("kernel_launches", C.c_int64),
        ("pass_ms_fused", C.c_double), ("pass_ms_grad", C.c_double), ("pass_ms_loss", C.c_double),
        ("grady_passes", C.c_int64), ("pass_ms_grady", C.c_double),
        ("i8_loss_passes", C.c_int64), ("i8_grad_passes", C.c_int64), ("i8_fallbacks", C.c_int64), ("i8_range", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Result(C.Structure):  # picard_result_t
    _fields_ = [
        ("n_components", C.c_int64), ("n_features", C.c_int64), ("n_samples", C.c_int64),
        ("whitening", dp), ("unmixing", dp), ("sources", dp), ("mean", dp),
        ("n_iterations", C.c_int64), ("converged", C.c_int32), ("gradient_norm", C.c_double), ("signs", dp),
        ("stats", Stats),
    ]


FLAG_NO_SPECULATION = 1
FLAG_KEEP_SOURCES_ON_DEVICE = 2
FLAG_NO_Y_STORE = 4
FLAG_FORCE_SPECULATION = 8
FLAG_NO_INT8 = 16
FLAG_FORCE_INT8 = 32
UNIQUE_ID_BYTES = 128

_lib = None


class LibraryMissing(ImportError):
    pass


def lib():
    """The loaded shared library. Raises LibraryMissing if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C picard-ica_b200/csrc`). picard_ica_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.picard_status_string.restype = C.c_char_p
        L.picard_result_free.restype = None
        L.picard_config_default.restype = None
        L.picard_core_destroy.restype = None
        L.picard_comm_destroy.restype = None
        L.picard_release_cache.restype = None
        _lib = L
    return _lib
