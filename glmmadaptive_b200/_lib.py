"""ctypes binding of the C ABI declared in include/glmma.h (libglmma.so).

This is the only way the Python host code reaches the engine; there is no CPU
fallback: if the shared library is missing, or no CUDA device is usable, the
calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglmma.so")
LIB_PATH = os.environ.get("GLMMA_LIB", LIB_PATH)      # development builds: make DEV=1 -> libglmma_dev.so

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NUMERIC = 0, 1, 2, 3, 4

FAMILY_IDS = {
    "binomial": 0, "poisson": 1, "negative binomial": 2, "zero-inflated poisson": 3,
    "zero-inflated negative binomial": 4, "zero-inflated binomial": 5, "hurdle poisson": 6,
    "hurdle negative binomial": 7, "hurdle log-normal": 8, "beta": 9, "hurdle beta": 10,
    "Gamma": 11, "censored normal": 12, "Student's-t": 13, "beta binomial": 14,
}
LINK_IDS = {None: 0, "default": 0, "logit": 1, "log": 2, "identity": 3, "probit": 4, "cloglog": 5}

c_double_p = C.POINTER(C.c_double)


class GlmmaData(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("link", C.c_int32),
        ("n_obs", C.c_int64), ("n_clusters", C.c_int64),
        ("p", C.c_int32), ("q_y", C.c_int32), ("p_zi", C.c_int32), ("q_zi", C.c_int32),
        ("n_phis", C.c_int32), ("y_cols", C.c_int32), ("nAGQ", C.c_int32), ("reserved", C.c_int32),
        ("y", c_double_p), ("X", c_double_p), ("Z", c_double_p), ("X_zi", c_double_p),
        ("Z_zi", c_double_p), ("offset", c_double_p), ("offset_zi", c_double_p),
        ("weights", c_double_p), ("id", C.POINTER(C.c_int32)), ("row_ptr", C.POINTER(C.c_int64)),
        ("family_df", C.c_double),
    ]


class GlmmaModeStats(C.Structure):
    _fields_ = [("n_not_converged", C.c_int64), ("n_chol_failed", C.c_int64),
                ("max_iterations", C.c_int32), ("reserved", C.c_int32),
                ("mean_iterations", C.c_double)]


# name -> (restype, argtypes); mirrors include/glmma.h one to one
_H = C.c_void_p
SIGNATURES = {
    "glmma_create": (C.c_int, [C.POINTER(GlmmaData), C.c_int, C.POINTER(_H)]),
    "glmma_create_multi": (C.c_int, [C.POINTER(GlmmaData), C.POINTER(C.c_int), C.c_int, C.POINTER(_H)]),
    "glmma_n_devices": (C.c_int, [_H]),
    "glmma_destroy": (None, [_H]),
    "glmma_last_error": (C.c_char_p, [_H]),
    "glmma_version": (C.c_int, []),
    "glmma_family_id": (C.c_int, [C.c_char_p]),
    "glmma_set_stream": (C.c_int, [_H, C.c_void_p]),
    "glmma_n_clusters": (C.c_int64, [_H]),
    "glmma_n_obs": (C.c_int64, [_H]),
    "glmma_nre": (C.c_int, [_H]),
    "glmma_n_nodes": (C.c_int, [_H]),
    "glmma_result_len": (C.c_int, [_H]),
    "glmma_find_modes": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int,
                                   C.POINTER(GlmmaModeStats)]),
    "glmma_get_modes": (C.c_int, [_H, c_double_p, c_double_p, c_double_p]),
    "glmma_set_modes": (C.c_int, [_H, c_double_p, c_double_p]),
    "glmma_eval": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p]),
    "glmma_eval_device": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_void_p]),
    "glmma_eval_contrib": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p,
                                     c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "glmma_em_estep": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "glmma_em_score": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p]),
    "glmma_glm_pass": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    "glmma_comm_export": (C.c_int, [_H, C.c_void_p]),
    "glmma_comm_connect": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "glmma_comm_size": (C.c_int, [_H]),
    "glmma_observed_information": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p]),
    "glmma_log_post_b": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "glmma_time_eval": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "glmma_measure_fp64_peak": (C.c_int, [C.c_int, c_double_p]),
    "glmma_launch_count": (C.c_int64, [_H]),
}

_lib = None


def load():
    """Loads libglmma.so (built by `make -C glmmadaptive_b200/csrc` or
    __graft_entry__.build()).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'); "
            "glmmadaptive_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GlmmaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"glmma error {code}: {msg}")
        self.code = code
