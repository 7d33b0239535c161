"""ctypes binding of libkb200.so (C ABI declared in include/kb200.h).

There is no CPU fallback: if the CUDA extension is missing or no GPU is visible, every
entry point raises.
"""
import ctypes
import os

import numpy as np

from . import _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_u64_p = ctypes.POINTER(ctypes.c_ulonglong)
c_void_p = ctypes.c_void_p

KERNEL_IDS = {"gaussian": 0, "iso_gaussian": 0, "exponential": 1, "matern32": 2, "matern52": 3}
FLAG_ISO, FLAG_NAIVE = 1, 2
OUT_PRED, OUT_SSQR, OUT_S, OUT_EI, OUT_POI, OUT_POF, OUT_UFUN = 1, 2, 4, 8, 16, 32, 64
NORM_NONE, NORM_DEFAULT, NORM_STD = 0, 1, 2
PROF_CLASSES = ["assemble", "chol_diag", "chol_panel", "finalize", "psi", "pred_solve", "epilogue", "other"]

# every symbol include/kb200.h declares: (restype, argtypes)
SIGNATURES = {
    "kb200_version": (ctypes.c_int, []),
    "kb200_last_error": (ctypes.c_char_p, []),
    "kb200_device_count": (ctypes.c_int, []),
    "kb200_host_alloc": (c_void_p, [ctypes.c_ulonglong]),
    "kb200_host_free": (None, [c_void_p]),
    "kb200_model_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int_p, ctypes.c_int, ctypes.c_double, ctypes.c_int]),
    "kb200_model_destroy": (None, [c_void_p]),
    "kb200_lik_batch": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, c_void_p, c_void_p]),
    "kb200_lik_batch_dev": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "kb200_fit_state": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       c_void_p, c_void_p, c_void_p, c_double_p, c_double_p, c_int_p,
                                       c_void_p, c_double_p, c_void_p]),
    "kb200_fit_append": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        c_void_p, c_void_p, c_void_p, c_double_p, c_double_p, c_int_p,
                                        c_void_p, c_double_p, c_void_p]),
    "kb200_predictor_load": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double]),
    "kb200_predict": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                     ctypes.c_uint, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_u64_p]),
    "kb200_predict_dev": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                         ctypes.c_uint, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_u64_p, c_void_p]),
    "kb200_model_dims": (ctypes.c_int, [c_void_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "kb200_launch_count": (ctypes.c_ulonglong, [c_void_p]),
    "kb200_profile_enable": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "kb200_profile_get": (ctypes.c_int, [c_void_p, c_double_p, c_u64_p, ctypes.c_int]),
    "kb200_fp64_probe": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p]),
    "kb200_set_groups": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "kb200_loocv": (ctypes.c_int, [c_void_p, c_void_p]),
    "kb200_akmcs_reduce": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                          c_double_p, ctypes.POINTER(ctypes.c_longlong), c_u64_p]),
    "kb200_sobol_sums": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                        c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "kb200_ehvi2d": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                    ctypes.c_longlong, ctypes.c_double, c_void_p]),
    "kb200_ehvi2d_dev": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_double, c_void_p,
                                        c_void_p]),
    "kb200_ehvi3d": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                    ctypes.c_longlong, ctypes.c_int, ctypes.c_double, c_void_p]),
    "kb200_ehvi3d_dev": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                        ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                        ctypes.c_double, c_void_p, c_void_p]),
    "kb200_debug_tiles": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p]),
    "kb200_debug_psi": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.c_int, c_void_p]),
}

_lib = None


class Kb200Error(RuntimeError):
    pass


def lib_path():
    # $KB200_LIB: an alternative build of the same sources (A/B tools only, e.g. tools/ab_predict_div.py)
    return os.environ.get("KB200_LIB") or _build.LIB


def load():
    """Load libkb200.so (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise Kb200Error(
            "kadal_b200: CUDA extension %s is missing — run `python -m kadal_b200._build` "
            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().kb200_last_error()
        raise Kb200Error("kb200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_gpu():
    lib = load()
    if lib.kb200_device_count() < 1:
        raise Kb200Error("kadal_b200: no CUDA device visible (there is no CPU fallback)")
    return lib


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by cudaHostAlloc'ed (pinned) memory."""
    lib = require_gpu()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.kb200_host_alloc(max(n, 8))
    if not p:
        raise Kb200Error("cudaHostAlloc failed")
    buf = (ctypes.c_char * max(n, 8)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.__array_interface__["data"][0], None)
    if p:
        load().kb200_host_free(p)
