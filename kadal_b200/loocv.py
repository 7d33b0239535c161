"""Leave-one-out cross validation: ``kadal/surrogate_models/supports/krigloocv2.py`` ('standard').

The reference refits the model n times on n-1 samples at the trained hyper-parameters and predicts
the left-out sample each time (O(n^4)).  Here the n predictions come from ONE pass over the
existing factor on the GPU (``kb200_loocv``: closed form, see kb200_kernels.cu).  The 'fast'
variant (krigloocv.py) inverts ``[[sigma^2 Psi, F], [F^T, ones]]`` -- note the block of ones where
the textbook formula has zeros, and Psi without the nugget -- which is a different quantity and not
on the GPU path; it raises instead of silently returning something else.
"""
import numpy as np

from .errperf import errperf
from .model import model_for
from .prediction import _check_supported, _ensure_predictor


def loocv(KrigInfo, errtype="rmse", type="standard"):
    """Returns ``(LOOCVerror, LOOCVpred[n, 1])`` like ``loocv2(KrigInfo, errtype)``."""
    if type == "fast":
        raise NotImplementedError(
            "kadal_b200: type='fast' (krigloocv.py, bordered inverse with a block of ones and no nugget) is "
            "not built; use type='standard' (krigloocv2.py), which is one GPU pass here")
    if type != "standard":
        raise ValueError('"type" only accept fast or standard')
    _check_supported(KrigInfo, None, None)
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    pred = m.loocv().reshape(-1, 1)
    y = np.asarray(KrigInfo["y"], dtype=float).reshape(-1, 1)
    return errperf(y, pred, errtype), pred
