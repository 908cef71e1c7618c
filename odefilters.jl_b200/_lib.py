"""ctypes binding of libpnde.so (include/pnde.h).  No CPU fallback: if the shared library is
missing the import of the compute entry points fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNDE_LIB") or os.path.join(_HERE, "libpnde.so")  # PNDE_LIB: experiment builds

ABI_VERSION = 2
MAX_DEVICES = 16
FLAG_REFERENCE_QUIRKS = 1
FLAG_ONE_THREAD = 2
ALG_EK0, ALG_EK1, ALG_IEKS = 0, 1, 2
DIFFUSIONS = {"dynamic": 0, "fixed": 1, "fixedMAP": 2, "dynamicMV": 3, "fixedMV": 4}
VF_KINDS = {"fhn_readme": 0, "fhn_lib": 1, "lotka_volterra": 2, "vanderpol": 3, "linear2": 4, "logistic": 5,
            "lorenz96": 6, "linear1": 7, "custom": 100}
VF_DIMS = {"fhn_readme": (2, 3), "fhn_lib": (2, 4), "lotka_volterra": (2, 4), "vanderpol": (2, 1), "linear2": (2, 2),
           "logistic": (1, 1), "linear1": (1, 1)}
SAVE_FINAL, SAVE_EVERY, SAVE_STRIDE = 0, 1, 2
HIST_FILTERED, HIST_SMOOTHED = 0, 1
RETCODES = {0: "Success", 1: "MaxIters", 2: "DtNaN", 3: "Unstable", 4: "HistoryFull", 5: "DtLessThanMin",
            6: "ZeroResidual"}


class PndeConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("alg", C.c_int32), ("order", C.c_int32), ("d", C.c_int32),
        ("vf_kind", C.c_int32), ("diffusion", C.c_int32), ("smooth", C.c_int32), ("adaptive", C.c_int32),
        ("save_mode", C.c_int32), ("save_stride", C.c_int32), ("device", C.c_int32), ("ieks_iterations", C.c_int32),
        ("abstol", C.c_double), ("reltol", C.c_double), ("dt", C.c_double), ("t0", C.c_double), ("t1", C.c_double),
        ("qmin", C.c_double), ("qmax", C.c_double), ("gamma", C.c_double), ("qsteady_min", C.c_double),
        ("qsteady_max", C.c_double), ("qoldinit", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double), ("maxiters", C.c_int64), ("max_saved", C.c_int64),
        ("n_devices", C.c_int32), ("flags", C.c_int32), ("device_list", C.c_int32 * MAX_DEVICES),
    ]


EXPORTS = [
    "pnde_default_config", "pnde_create", "pnde_create_custom", "pnde_check_custom", "pnde_destroy", "pnde_last_error", "pnde_state_dim", "pnde_n_params",
    "pnde_record_len", "pnde_cov_len", "pnde_solve_ensemble", "pnde_solve_ensemble_to_host", "pnde_upload", "pnde_run", "pnde_synchronize", "pnde_last_run_ms",
    "pnde_last_launch_count", "pnde_smooth", "pnde_query_sizes", "pnde_get_counts", "pnde_get_final",
    "pnde_get_history", "pnde_get_history_sqrt", "pnde_step_from_state", "pnde_get_marginals", "pnde_sample", "pnde_dense_sample", "pnde_eval_dense", "pnde_measure_fp64_peak",
    "pnde_measure_hbm_copy", "pnde_host_alloc", "pnde_host_free",
]

_lib = None


def load():
    """Load libpnde.so; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C odefilters.jl_b200/csrc` "
            "(there is no CPU fallback for the ODE-filter hot path)")
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip64, ip32 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.pnde_default_config.argtypes = [C.POINTER(PndeConfig), C.c_int32, C.c_int32, C.c_int32]
    lib.pnde_create.argtypes = [C.POINTER(PndeConfig), C.POINTER(vp)]
    lib.pnde_create_custom.argtypes = [C.POINTER(PndeConfig), C.c_int32, C.c_int32, C.c_char_p, C.c_char_p,
                                       C.POINTER(vp)]
    lib.pnde_check_custom.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_int64]
    lib.pnde_destroy.argtypes = [vp]
    lib.pnde_last_error.argtypes = [vp]
    lib.pnde_last_error.restype = C.c_char_p
    for f in ("pnde_state_dim", "pnde_n_params", "pnde_record_len", "pnde_cov_len", "pnde_last_launch_count"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = C.c_int64
    lib.pnde_solve_ensemble.argtypes = [vp, C.c_int64, vp, vp]
    lib.pnde_solve_ensemble_to_host.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp]
    lib.pnde_upload.argtypes = [vp, C.c_int64, vp, vp]
    lib.pnde_run.argtypes = [vp]
    lib.pnde_synchronize.argtypes = [vp]
    lib.pnde_smooth.argtypes = [vp]
    lib.pnde_last_run_ms.argtypes = [vp, dp, dp]
    lib.pnde_query_sizes.argtypes = [vp, ip64, ip64]
    lib.pnde_get_counts.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pnde_get_final.argtypes = [vp, vp, vp, vp, vp]
    lib.pnde_get_history.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, vp, vp, vp, vp, vp]
    lib.pnde_get_history_sqrt.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, vp, vp]
    lib.pnde_step_from_state.argtypes = [vp, C.c_int64] + [vp] * 13
    lib.pnde_get_marginals.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, vp, vp, vp, vp]
    lib.pnde_sample.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.c_uint64, vp, vp, vp]
    lib.pnde_dense_sample.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int32, C.c_uint64, vp]
    lib.pnde_eval_dense.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp]
    lib.pnde_measure_fp64_peak.argtypes = [C.c_int32, dp]
    lib.pnde_measure_hbm_copy.argtypes = [C.c_int32, dp]
    lib.pnde_host_alloc.argtypes = [C.POINTER(vp), C.c_int64]
    lib.pnde_host_free.argtypes = [vp]
    _lib = lib
    return lib
