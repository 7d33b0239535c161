"""Leave-one-out cross validation: ``kadal/surrogate_models/supports/krigloocv2.py`` ('standard').

The reference refits the model n times on n-1 samples at the trained hyper-parameters and predicts
the left-out sample each time (O(n^4)).  Here the n predictions come from ONE pass over the
existing factor on the GPU (``kb200_loocv``: closed form, see kb200_kernels.cu).  The 'fast'
variant (krigloocv.py) inverts ``[[sigma^2 Psi, F], [F^T, ones]]`` -- note the block of ones where
the textbook formula has zeros, and ``KrigInfo['Psi']``, i.e. Psi WITHOUT the nugget -- a different
quantity.  It is evaluated in closed form as well (``kb200_loocv_fast``), from a Cholesky factor of Psi
itself: a second, throw-away device model of the same samples is fitted at the trained hyper-parameters
with a zero nugget.  Where that Psi is not positive definite to working precision (the reference's general
LU inverse returns noise there) ``numpy.linalg.LinAlgError`` is raised.
"""
import numpy as np

from .errperf import errperf
from .model import model_for
from .prediction import _check_supported, _ensure_predictor


def loocv(KrigInfo, errtype="rmse", type="standard"):
    """Returns ``(LOOCVerror, LOOCVpred[n, 1])`` like ``loocv2(KrigInfo, errtype)``."""
    if type not in ("fast", "standard"):
        raise ValueError('"type" only accept fast or standard')
    _check_supported(KrigInfo, None, None)
    if type == "fast":
        return _loocv_fast(KrigInfo, errtype)
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    pred = m.loocv().reshape(-1, 1)
    y = np.asarray(KrigInfo["y"], dtype=float).reshape(-1, 1)
    return errperf(y, pred, errtype), pred


def _loocv_fast(KrigInfo, errtype):
    """krigloocv.py:4-38 on the GPU (see the module docstring)."""
    from .model import GpuModel, kernel_ids, select_xy
    if np.shape(KrigInfo["F"])[1] != 1:
        raise NotImplementedError("kadal_b200: 'fast' LOOCV is built for ordinary Kriging (TrendOrder 0)")
    X, y = select_xy(KrigInfo)
    kids, iso = kernel_ids(KrigInfo)
    kpls = str(KrigInfo["type"]).lower() == "kpls"
    m = GpuModel(X, y, KrigInfo["F"], kids, pls=KrigInfo["plscoeff"] if kpls else None, nsamp=KrigInfo["nsamp"], iso=iso)
    try:
        theta = np.asarray(KrigInfo["Theta"], dtype=float).ravel()
        x = theta[:1] if iso else theta
        if len(kids) > 1:
            x = np.r_[x, np.asarray(KrigInfo["wgkf"], dtype=float).ravel()]
        st = m.fit_state(x, False, -300.0, want_U=False, want_Psi=False)      # eps = 1e-300: Psi itself
        if st["info"] != 0:
            raise np.linalg.LinAlgError("Psi without the nugget is not positive definite to working precision")
        s2 = float(np.asarray(KrigInfo["SigmaSqr"]).ravel()[0])
        pred = m.loocv_fast(s2).reshape(-1, 1)
    finally:
        m.close()
    yv = np.asarray(KrigInfo["y"], dtype=float).reshape(-1, 1)
    return errperf(yv, pred, errtype), pred


def loocv2(KrigInfo, errtype="rmse"):
    """Drop-in for ``krigloocv2.loocv2(KrigInfo, errtype)`` (bound in kriging_model.py:8): the closed form."""
    return loocv(KrigInfo, errtype=errtype, type="standard")


def loocv_fast(KrigInfo, errtype="rmse", num=None):
    """Drop-in for ``krigloocv.loocv(KrigInfo, errtype, num)`` (bound in kriging_model.py:7)."""
    if num is not None:
        raise NotImplementedError("kadal_b200: num= (multi-objective list layout of KrigInfo) is not supported")
    return loocv(KrigInfo, errtype=errtype, type="fast")
