"""Device-resident view of a KrigInfo dict.

The KrigInfo dict stays the single source of truth and stays picklable: device handles live
in a side cache keyed by the identity (+ cheap checksum) of the arrays the reference reads
(``likelihood_func.py:38-68``, ``prediction.py:122-157``), never inside the dict.
"""
import ctypes
import os
import zlib
from collections import OrderedDict

import numpy as np

from . import _capi
from ._capi import c_void_p, ptr, as_f64, check


def select_xy(KrigInfo):
    """Same selection as likelihood_func.py:44-53 / prediction.py:138-146."""
    if KrigInfo["standardization"] is False:
        return KrigInfo["X"], KrigInfo["y"]
    X = KrigInfo["X_norm"]
    if "y_norm" in KrigInfo and np.any(np.asarray(KrigInfo["y_norm"]).ravel()[0] != 0):
        return X, KrigInfo["y_norm"]
    return X, KrigInfo["y"]


def kernel_ids(KrigInfo):
    kern = KrigInfo["kernel"]
    if not isinstance(kern, (list, tuple)):
        kern = [kern]
    ids = []
    for k in kern:
        try:
            ids.append(_capi.KERNEL_IDS[k.lower()])
        except KeyError:
            raise ValueError("Kernel option not available!")  # kernelfunc.py:31
    return ids, list(kern) == ["iso_gaussian"]


class GpuModel:
    """Owns one ``kb200_model`` handle (training data + workspaces + predictor state)."""

    def __init__(self, X, y, F, kernels, pls=None, nsamp=None, iso=False, device=None, naive=None):
        lib = _capi.require_gpu()
        self.lib = lib
        X = as_f64(X)
        y = as_f64(np.asarray(y).reshape(-1))
        F = as_f64(F)
        if F.ndim == 1:
            F = F.reshape(-1, 1)
        self.n, self.d = X.shape
        self.p = F.shape[1]
        if y.shape[0] != self.n or F.shape[0] != self.n:
            raise ValueError("X, y, F row counts differ")
        self.h = 0
        if pls is not None:
            pls = as_f64(pls)
            self.h = pls.shape[1]
        self.ntheta = self.h if self.h else self.d
        self.nkernel = len(kernels)
        self.kernels = tuple(int(k) for k in kernels)
        self.iso = bool(iso)
        if device is None:
            device = int(os.environ.get("KADAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            device = device % max(1, lib.kb200_device_count())
        if naive is None:
            naive = os.environ.get("KB200_NAIVE", "0") == "1"
        flags = (_capi.FLAG_ISO if iso else 0) | (_capi.FLAG_NAIVE if naive else 0)
        kid = (ctypes.c_int * self.nkernel)(*kernels)
        h = c_void_p()
        check(lib.kb200_model_create(ctypes.byref(h), device, self.n, self.d, self.p, self.h, ptr(X), ptr(y),
                                     ptr(F), ptr(pls), kid, self.nkernel,
                                     float(self.n if nsamp is None else nsamp), flags))
        self._h = h
        self.device = device
        self.ymin = float(np.min(y))
        self.pred_token = None
        self._keep = None
        self.fit_key = None       # (x bytes, trainvar, nugget) of the fit an append may continue from

    def close(self):
        if getattr(self, "_h", None):
            self.lib.kb200_model_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        """The live ``kb200_model*``; a model evicted from the cache (or closed by hand) raises instead of handing
        a dangling pointer to the C library."""
        h = getattr(self, "_h", None)
        if not h:
            raise _capi.Kb200Error("GpuModel is closed (evicted from the model cache: raise $KADAL_B200_CACHE, or "
                                   "obtain the model again with model_for(KrigInfo))")
        return h

    @property
    def closed(self):
        return not getattr(self, "_h", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ likelihood
    def lik_batch(self, Xpop, trainvar, nugget_log10):
        Xpop = as_f64(np.atleast_2d(Xpop))
        P, nb = Xpop.shape
        nll = np.empty(P)
        info = np.empty(P, dtype=np.int32)
        check(self.lib.kb200_lik_batch(self.handle, ptr(Xpop), P, nb, int(bool(trainvar)),
                                       float(nugget_log10), ptr(nll), ptr(info)))
        return nll, info

    def lik_batch_dev(self, x_ptr, P, nbhyp, trainvar, nugget_log10, nll_ptr, info_ptr, stream=0):
        """Device-pointer batch.  ``stream`` is a raw ``cudaStream_t`` (e.g. ``torch.cuda.current_stream().cuda_stream``):
        the call is asynchronous and ordered on that stream.  ``stream=0`` runs on the model's own stream and BLOCKS
        until the results are complete.  One call in flight per model: successive calls on different streams are
        serialised by the library (include/kb200.h)."""
        check(self.lib.kb200_lik_batch_dev(self.handle, c_void_p(x_ptr), P, nbhyp, int(bool(trainvar)),
                                           float(nugget_log10), c_void_p(nll_ptr), c_void_p(info_ptr),
                                           c_void_p(stream)))

    def fit_state(self, x, trainvar, nugget_log10, want_U=True, want_Psi=True, append_to=None):
        """likelihood(x, KrigInfo, 'all') on the device.  ``append_to``: a GpuModel holding the fit of the same
        ``x`` on this model's samples minus the last one -- the factor is then extended by one row
        (``kb200_fit_append``) instead of being recomputed."""
        x = as_f64(np.asarray(x).reshape(-1))
        n = self.n
        U = _capi.result_empty((n, n)) if want_U else None
        Psi = _capi.result_empty((n, n)) if want_Psi else None
        BE = np.empty(self.p)
        theta = np.empty(self.ntheta)
        wgkf = np.empty(self.nkernel)
        s2, nll, nug = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        info = ctypes.c_int()
        tail = (ptr(x), x.shape[0], int(bool(trainvar)), float(nugget_log10),
                ptr(U), ptr(Psi), ptr(BE), ctypes.byref(s2), ctypes.byref(nll),
                ctypes.byref(info), ptr(theta), ctypes.byref(nug), ptr(wgkf))
        if append_to is not None:
            check(self.lib.kb200_fit_append(self.handle, append_to.handle, *tail))
        else:
            check(self.lib.kb200_fit_state(self.handle, *tail))
        self.fit_key = (x.tobytes(), bool(trainvar), float(nugget_log10)) if info.value == 0 and self.p == 1 else None
        self.last_fit_appended = append_to is not None
        return dict(U=U, Psi=Psi, BE=BE, sigma2=s2.value, nll=nll.value, info=info.value, theta=theta,
                    nugget=nug.value, wgkf=wgkf)

    def debug_psi(self, x, trainvar, nugget_log10, add_eps=False):
        x = as_f64(np.asarray(x).reshape(-1))
        Psi = np.empty((self.n, self.n))
        check(self.lib.kb200_debug_psi(self.handle, ptr(x), x.shape[0], int(bool(trainvar)), float(nugget_log10),
                                       int(add_eps), ptr(Psi)))
        return Psi

    # ------------------------------------------------------------------ prediction
    def predictor_load(self, theta_log10, wgkf, U, BE, sigma2):
        self.fit_key = None
        theta_log10 = as_f64(np.asarray(theta_log10).reshape(-1))
        wgkf = as_f64(np.asarray(wgkf, dtype=float).reshape(-1))
        U = as_f64(U)
        BE = as_f64(np.asarray(BE).reshape(-1))
        check(self.lib.kb200_predictor_load(self.handle, ptr(theta_log10), ptr(wgkf), ptr(U), ptr(BE),
                                            float(sigma2)))

    def predict(self, x, want, normtype=0, norm_a=None, norm_b=None, ymin=None, limit=0.0, limit_ge=True,
                out=None):
        """Host-pointer predict: returns dict name -> (npred,) array plus 'stats'."""
        x = as_f64(np.atleast_2d(x))
        npred = x.shape[0]
        if x.shape[1] != self.d:
            raise ValueError("prediction sites have %d columns, model has %d" % (x.shape[1], self.d))
        names = ["pred", "ssqr", "s", "ei", "poi", "pof", "ufun"]
        outs = {}
        ptrs = []
        for i, nm in enumerate(names):
            if want & (1 << i):
                arr = out[nm] if out is not None and nm in out else _capi.result_empty((npred,))
                outs[nm] = arr
                ptrs.append(ptr(arr))
            else:
                ptrs.append(None)
        stats = (ctypes.c_ulonglong * 2)()
        na = as_f64(norm_a) if norm_a is not None else None
        nb = as_f64(norm_b) if norm_b is not None else None
        check(self.lib.kb200_predict(self.handle, ptr(x), npred, int(normtype), ptr(na), ptr(nb), int(want),
                                     float(self.ymin if ymin is None else ymin), float(limit), int(bool(limit_ge)),
                                     *ptrs, stats))
        outs["stats"] = (int(stats[0]), int(stats[1]))
        return outs

    def predict_dev(self, x_ptr, npred, want, out_ptrs, normtype=0, norm_a=None, norm_b=None, ymin=None,
                    limit=0.0, limit_ge=True, stream=0, want_stats=False):
        """Device-pointer predict (x_ptr and out_ptrs are raw device addresses, e.g. torch
        ``tensor.data_ptr()``); no host<->device traffic except the tiny normalisation vectors."""
        na = as_f64(norm_a) if norm_a is not None else None
        nb = as_f64(norm_b) if norm_b is not None else None
        ptrs = [c_void_p(p) if p else None for p in out_ptrs]
        stats = (ctypes.c_ulonglong * 2)() if want_stats else None
        check(self.lib.kb200_predict_dev(self.handle, c_void_p(x_ptr), int(npred), int(normtype), ptr(na), ptr(nb),
                                         int(want), float(self.ymin if ymin is None else ymin), float(limit),
                                         int(bool(limit_ge)), *ptrs, stats, c_void_p(stream)))
        return (int(stats[0]), int(stats[1])) if want_stats else None

    def akmcs_reduce(self, x, normtype=0, norm_a=None, norm_b=None):
        """(min U, argmin U, [#(G<=0), #(G-1.96s<=0), #(G+1.96s<=0)]) over the population x."""
        x = as_f64(np.atleast_2d(x))
        if x.shape[1] != self.d:
            raise ValueError("population has %d columns, model has %d" % (x.shape[1], self.d))
        na = as_f64(norm_a) if norm_a is not None else None
        nb = as_f64(norm_b) if norm_b is not None else None
        mu, arg, cnt = ctypes.c_double(), ctypes.c_longlong(), (ctypes.c_ulonglong * 3)()
        check(self.lib.kb200_akmcs_reduce(self.handle, ptr(x), x.shape[0], int(normtype), ptr(na), ptr(nb),
                                          ctypes.byref(mu), ctypes.byref(arg), cnt))
        return float(mu.value), int(arg.value), [int(c) for c in cnt]

    def akmcs_reduce_dev(self, x_ptr, npred, normtype=0, norm_a=None, norm_b=None, stream=0):
        """The same over a population that is already resident in HBM (``x_ptr``: raw device address of
        [npred][d] doubles, e.g. ``tensor.data_ptr()``): nothing but the scalars crosses PCIe."""
        na = as_f64(norm_a) if norm_a is not None else None
        nb = as_f64(norm_b) if norm_b is not None else None
        mu, arg, cnt = ctypes.c_double(), ctypes.c_longlong(), (ctypes.c_ulonglong * 3)()
        check(self.lib.kb200_akmcs_reduce_dev(self.handle, c_void_p(x_ptr), int(npred), int(normtype), ptr(na), ptr(nb),
                                              ctypes.byref(mu), ctypes.byref(arg), cnt, c_void_p(stream)))
        return float(mu.value), int(arg.value), [int(c) for c in cnt]

    def loocv(self):
        """Leave-one-out predictions (n,) of the installed fit (krigloocv2.py semantics)."""
        out = np.empty(self.n, dtype=np.float64)
        check(self.lib.kb200_loocv(self.handle, ptr(out)))
        return out

    def loocv_fast(self, sigma2):
        """krigloocv.py's 'fast' leave-one-out values (n,) from the installed fit of Psi WITHOUT nugget (kb200.h)."""
        out = np.empty(self.n, dtype=np.float64)
        check(self.lib.kb200_loocv_fast(self.handle, float(sigma2), ptr(out)))
        return out

    def set_groups(self, ngroups):
        """Number of candidate groups (streams) of a likelihood batch; 1 serialises the kernels."""
        check(self.lib.kb200_set_groups(self.handle, int(ngroups)))

    def set_schedule_population(self, population):
        """Choose the factorisation schedule as for a population of this size (0: from each call's own P); what a
        rank evaluating a shard passes so that its bits equal the single-GPU result (kb200.h)."""
        check(self.lib.kb200_set_schedule_population(self.handle, int(population)))

    def profile(self, on):
        check(self.lib.kb200_profile_enable(self.handle, int(bool(on))))

    def profile_get(self, reset=True):
        ms = (ctypes.c_double * 8)()
        cnt = (ctypes.c_ulonglong * 8)()
        check(self.lib.kb200_profile_get(self.handle, ms, cnt, int(reset)))
        return {nm: (ms[i], int(cnt[i])) for i, nm in enumerate(_capi.PROF_CLASSES)}

    def launch_count(self):
        return int(self.lib.kb200_launch_count(self.handle))


# ------------------------------------------------------------------------- cache
_CACHE = OrderedDict()
_CACHE_MAX = int(os.environ.get("KADAL_B200_CACHE", "8"))


try:                                   # XXH3 when the image has it (0.17 ms at C5's 3.2 MB, like the float sum it replaces);
    from xxhash import xxh3_64_intdigest as _hash_bytes
except Exception:                      # CRC-32 otherwise (0.72 ms)
    _hash_bytes = zlib.crc32


def _fingerprint(a):
    """Identity + content of an array the device copy was made from: an in-place edit of any kind (a swapped pair of
    rows keeps the sum) must give a new key.  A hash of the bytes, O(size) like the sum it replaces."""
    a = np.asarray(a)
    c = a if a.flags.c_contiguous else np.ascontiguousarray(a)
    return (id(a), a.shape, str(a.dtype), _hash_bytes(memoryview(c).cast("B")) if a.size else 0)


def _content(a):
    return _fingerprint(a)[1:]


def model_for(KrigInfo):
    """GpuModel for this KrigInfo (created on first use, re-created when the training data
    the reference would read has changed, e.g. after SOBO/AK-MCS append a sample)."""
    X, y = select_xy(KrigInfo)
    F = KrigInfo["F"]
    kids, iso = kernel_ids(KrigInfo)
    kpls = str(KrigInfo["type"]).lower() == "kpls"
    pls = KrigInfo["plscoeff"] if kpls else None
    if not kpls and str(KrigInfo["type"]).lower() != "kriging":
        raise ValueError("Kriging model dictionary 'type' value: '%s' is not recognised." % KrigInfo["type"])
    key = (_fingerprint(X), _fingerprint(y), _fingerprint(F), tuple(kids), iso,
           _fingerprint(pls) if pls is not None else None, KrigInfo["nsamp"])
    m = _CACHE.get(key)
    if m is not None:
        _CACHE.move_to_end(key)
        return m
    if kpls and KrigInfo["n_princomp"] is not False and int(KrigInfo["n_princomp"]) != np.shape(pls)[1]:
        raise ValueError("plscoeff has %d components, n_princomp=%s" % (np.shape(pls)[1], KrigInfo["n_princomp"]))
    m = GpuModel(X, y, F, kids, pls=pls, nsamp=KrigInfo["nsamp"], iso=iso)
    m._keep = (X, y, F, pls)   # keep the arrays alive so ids stay unique
    m._uploaded = tuple(_content(a) if a is not None else None for a in m._keep)   # what the device copy holds
    _CACHE[key] = m
    while len(_CACHE) > _CACHE_MAX:
        _, old = _CACHE.popitem(last=False)
        old.close()
    return m


def append_source(m, x, trainvar, nugget_log10):
    """A cached model that ``m`` extends by exactly one sample and that holds the fit of the same
    hyper-parameters (Kriging believer, MOBO.py:344-356; AK-MCS / SOBO sequential updates), or None.
    ``$KB200_NO_APPEND=1`` always answers None (every 'all' call refactorises)."""
    if os.environ.get("KB200_NO_APPEND", "0") == "1" or m.p != 1 or m._keep is None or m.n < 2:
        return None
    x = as_f64(np.asarray(x).reshape(-1))
    want = (x.tobytes(), bool(trainvar), float(nugget_log10))
    X, y, F, pls = m._keep
    for prev in reversed(list(_CACHE.values())):
        if prev is m or prev.closed or prev.fit_key != want or prev._keep is None:
            continue
        # prev's host arrays must still be what was uploaded (an in-place edit after the upload would let the
        # append extend a factor of other data: the C side leaves the prefix check to the caller)
        if getattr(prev, "_uploaded", None) != tuple(_content(a) if a is not None else None for a in prev._keep):
            continue
        if (prev.n != m.n - 1 or prev.d != m.d or prev.p != m.p or prev.h != m.h or prev.iso != m.iso
                or prev.device != m.device or prev.kernels != m.kernels):
            continue
        pX, py, pF, ppls = prev._keep
        n = prev.n
        if not (np.array_equal(np.asarray(X)[:n], pX) and np.array_equal(np.asarray(y).reshape(-1)[:n], np.asarray(py).reshape(-1))
                and np.array_equal(np.asarray(F)[:n], pF)):
            continue
        if (pls is None) != (ppls is None) or (pls is not None and not np.array_equal(pls, ppls)):
            continue
        return prev
    return None


def clear_cache():
    while _CACHE:
        _, m = _CACHE.popitem()
        m.close()


def fp64_probe(device=0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) measured with register-resident loops."""
    lib = _capi.require_gpu()
    out = []
    for mode in (0, 1):
        v = ctypes.c_double()
        check(lib.kb200_fp64_probe(device, mode, ctypes.byref(v)))
        out.append(v.value)
    return tuple(out)
