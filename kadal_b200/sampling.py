"""Host-side mirror of the two ``kadal/misc/sampling/samplingplan.py`` entry points the
Kriging hot path uses: ``standardize`` (:53-103) and ``sampling`` (:13-35, Sobol start
points of the hyper-parameter search).  Elementwise O(n d) host work, as in the reference.
"""
import numpy as np


def standardize(X, y=None, type="default", norm_y=False, **kwargs):
    """Same contract as samplingplan.standardize (:53-103).

    'default': ((X - lo)/(hi - lo) - 0.5)*2 with ``range=[lo; hi]`` (2 x nvar, plus the y
    columns when ``norm_y``); 'std': (X - mean)/std(ddof=1), zero spreads replaced by 1.
    Return tuples match the reference: X_norm | (X_norm, y_norm) | (X_norm, X_mean, X_std)
    | (X_norm, y_norm, X_mean, y_mean, X_std, y_std)."""
    X = np.asarray(X, dtype=float)
    kind = type.lower()
    if kind == "default":
        ranges = kwargs.get("range", None)
        if ranges is None or np.asarray(ranges).dtype == object:
            raise ValueError("Default normalization requires range value!")
        ranges = np.asarray(ranges, dtype=float)
        nvar = X.shape[1]
        X_norm = np.empty(X.shape)
        for i in range(nvar):
            X_norm[:, i] = (X[:, i] - ranges[0, i]) / (ranges[1, i] - ranges[0, i])
        X_norm = (X_norm - 0.5) * 2
        if not norm_y:
            return X_norm
        y = np.asarray(y, dtype=float)
        y_norm = np.empty(y.shape)
        for j in range(y.shape[1]):
            lo, hi = ranges[0, nvar + j], ranges[1, nvar + j]
            y_norm[:, j] = (y[:, j] - lo) / (hi - lo)
        return X_norm, (y_norm - 0.5) * 2
    if kind == "std":
        X_mean = np.mean(X, axis=0)
        X_std = X.std(axis=0, ddof=1)
        X_std[X_std == 0.0] = 1.0
        X_norm = (X - X_mean) / X_std
        if not norm_y:
            return X_norm, X_mean, X_std
        y = np.asarray(y, dtype=float)
        y_mean = np.mean(y, axis=0)
        y_std = y.std(axis=0, ddof=1)
        y_std[y_std == 0.0] = 1.0
        return X_norm, (y - y_mean) / y_std, X_mean, y_mean, X_std, y_std
    raise ValueError("standardization type '%s' is not recognised" % type)


def realval(lb, ub, samp):
    """samplingplan.realval (:37-51): scale [0,1] samples to [lb, ub]."""
    lb = np.asarray(lb, dtype=float)
    ub = np.asarray(ub, dtype=float)
    samp = np.asarray(samp, dtype=float)
    if len(ub) != len(lb):
        raise ValueError("Lower and upper bound have to be in the same dimension")
    if len(ub) != samp.shape[1]:
        raise ValueError("sample and bound are not in the same dimension")
    return samp * (ub - lb) + lb


def sobol_unit(nsamp, nvar):
    """Rows 1..nsamp of the unscrambled Sobol sequence (the reference drops the all-zero
    first row, samplingplan.py:18-19).  Uses the ``sobolsampling`` package when it is
    installed (what KADAL itself calls), otherwise scipy's generator; the two differ in
    direction numbers, which only moves the optimiser's start points."""
    try:
        from sobolsampling.sobol_new import sobol_points
        return np.asarray(sobol_points(nsamp + 1, nvar))[1:, :]
    except Exception:
        from scipy.stats import qmc
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return qmc.Sobol(nvar, scramble=False).random(nsamp + 1)[1:, :]


def sobol_direction_numbers(nvar, nsamp):
    """(v, bits): the direction numbers ``v[nvar][bits]`` (uint64) of the generator :func:`sobol_unit` uses, for the
    device-side generator (``kb200_sobol_sums_gen`` / ``kb200_sobol_points_dev``), which reproduces
    ``sobol_unit(nsamp, nvar)`` bit for bit without any host array: point ``i`` is the xor of ``v[:, b]`` over the set
    bits ``b`` of ``i ^ (i >> 1)``, times ``2**-bits``.  ``None`` when the host generator is the ``sobolsampling``
    package (its tables are not exposed; callers then generate on the host and upload)."""
    try:
        import sobolsampling.sobol_new  # noqa: F401
        return None
    except Exception:
        pass
    try:
        from scipy.stats import qmc
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            eng = qmc.Sobol(nvar, scramble=False)
        bits = int(eng.bits)
        v = np.ascontiguousarray(eng._sv, dtype=np.uint64)      # (nvar, bits) direction numbers of scipy's generator
    except Exception:                                           # another scipy layout: generate on the host instead
        return None
    if v.shape != (nvar, bits) or nsamp + 1 > (1 << bits):
        return None
    return v, bits


def sobol_points_device(nsamp, nvar, lobound=None, upbound=None, first_row=1, device=None, out_ptr=None, stream=0):
    """Rows ``first_row .. first_row + nsamp - 1`` of the plan ``sampling('sobol', nvar, ...)`` draws (``first_row=1``:
    the reference drops row 0), written to device memory at ``out_ptr`` ([nsamp][nvar] doubles), scaled like
    ``realval`` when bounds are given.  Returns False when the host generator cannot be reproduced on the device."""
    import ctypes
    import os
    from . import _capi
    dn = sobol_direction_numbers(nvar, first_row + nsamp)
    if dn is None:
        return False
    v, bits = dn
    lib = _capi.require_gpu()
    if device is None:
        device = int(os.environ.get("KADAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0"))) % max(1, lib.kb200_device_count())
    lo = _capi.as_f64(lobound) if lobound is not None else None
    hi = _capi.as_f64(upbound) if upbound is not None else None
    _capi.check(lib.kb200_sobol_points_dev(int(device), v.ctypes.data_as(ctypes.c_void_p), bits, int(nvar), int(first_row),
                                           int(nsamp), _capi.ptr(lo), _capi.ptr(hi), ctypes.c_void_p(out_ptr),
                                           ctypes.c_void_p(stream)))
    return True


def sampling(option, nvar, nsamp, result="normalized", upbound=None, lobound=None):
    """samplingplan.sampling (:13-35) for the options the hyper-parameter search uses."""
    opt = option.lower()
    if opt in ("sobol", "sobolnew"):
        samplenorm = sobol_unit(nsamp, nvar)
    elif opt == "random":
        samplenorm = np.random.rand(nsamp, nvar)
    else:
        raise NameError("sampling plan unavailable!")
    if result.lower() == "real":
        if lobound is None or upbound is None:
            raise ValueError("lb and ub must have value")
        if np.any(np.asarray(upbound) - np.asarray(lobound) < 0):
            raise ValueError("Upper bound must bigger than lower bound!")
        return samplenorm, realval(lobound, upbound, samplenorm)
    return samplenorm, np.zeros(np.shape(samplenorm))
