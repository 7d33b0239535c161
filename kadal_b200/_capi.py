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
    "kb200_set_schedule_population": (ctypes.c_int, [c_void_p, ctypes.c_longlong]),
    "kb200_loocv": (ctypes.c_int, [c_void_p, c_void_p]),
    "kb200_loocv_fast": (ctypes.c_int, [c_void_p, ctypes.c_double, c_void_p]),
    "kb200_akmcs_reduce": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                          c_double_p, ctypes.POINTER(ctypes.c_longlong), c_u64_p]),
    "kb200_akmcs_reduce_dev": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                              c_double_p, ctypes.POINTER(ctypes.c_longlong), c_u64_p, c_void_p]),
    "kb200_sobol_sums_dev": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p, c_void_p,
                                            c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kb200_sobol_sums_gen": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                                            c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p, ctypes.c_int,
                                            c_void_p, c_void_p, c_void_p]),
    "kb200_sobol_points_dev": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                              ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
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
    "kb200_debug_exp": (ctypes.c_int, [ctypes.c_int, c_void_p, ctypes.c_longlong, ctypes.c_int, c_void_p]),
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


# ---------------------------------------------------------------- pinned result arrays
# likelihood(x, KrigInfo, 'all') hands U and Psi (n x n each) back to the host: into pageable numpy memory that copy
# is 5 of the 6 ms of the call at n = 1000 (56 of 60 ms at n = 4000).  Result arrays therefore live in recycled pinned
# buffers: a buffer goes back to the free list when the last numpy view of it dies, pinning new pages (slow) is
# needed only while the working set grows.  Bounded: beyond $KB200_PINNED_MB (default 2048) of live + free pinned
# memory results fall back to ordinary numpy arrays (same values, slower copy).
import threading
import weakref

_RES_LOCK = threading.Lock()
_RES_FREE = []            # (capacity, pointer)
_RES_BYTES = [0]          # live + free
_RES_LIMIT = int(float(os.environ.get("KB200_PINNED_MB", "2048")) * (1 << 20))


def _res_release(p, cap):
    with _RES_LOCK:
        if len(_RES_FREE) < 16:
            _RES_FREE.append((cap, p))
            return
        _RES_BYTES[0] -= cap
    try:
        load().kb200_host_free(c_void_p(p))
    except Exception:
        pass


def result_empty(shape):
    """float64 array for a device->host result, pinned when the budget allows."""
    count = int(np.prod(shape))
    need = max(count * 8, 8)
    if _RES_LIMIT <= 0 or need < (1 << 16):
        return np.empty(shape)
    p, drop = None, []
    with _RES_LOCK:
        fits = [i for i, (c, _) in enumerate(_RES_FREE) if need <= c <= 2 * need]
        if fits:
            cap, p = _RES_FREE.pop(min(fits, key=lambda i: _RES_FREE[i][0]))
        else:
            cap = need + need // 8
            if _RES_BYTES[0] + cap > _RES_LIMIT:      # make room: the free buffers did not fit anyway
                drop, _RES_FREE[:] = list(_RES_FREE), []
                _RES_BYTES[0] -= sum(c for c, _ in drop)
            if _RES_BYTES[0] + cap > _RES_LIMIT:
                cap = None
            else:
                _RES_BYTES[0] += cap
    lib = require_gpu()
    for _, q in drop:
        lib.kb200_host_free(c_void_p(q))
    if cap is None:
        return np.empty(shape)
    if p is None:
        p = lib.kb200_host_alloc(cap)
        if not p:
            with _RES_LOCK:
                _RES_BYTES[0] -= cap
            return np.empty(shape)
    buf = (ctypes.c_char * cap).from_address(p)
    weakref.finalize(buf, _res_release, p, cap).atexit = False
    return np.frombuffer(buf, dtype=np.float64, count=count).reshape(shape)
