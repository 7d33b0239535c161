"""Drop-in for ``kadal/surrogate_models/supports/prediction.py``.

``prediction(x, KrigInfo, predtypes)`` keeps the reference's signature, output shapes and
error types (prediction.py:71-323); ``predict_arrays`` is the additive bulk entry point
(flat outputs, any ``npred`` in one call, optional U-function).  The arithmetic —
standardisation, cross-correlation, mean, variance, acquisition values — runs in the CUDA
library (``kb200_predict``).
"""
import numpy as np

from . import _capi
from .likelihood_func import _pred_token
from .model import model_for

_WANT = {"pred": _capi.OUT_PRED, "ssqr": _capi.OUT_SSQR, "s": _capi.OUT_S, "ei": _capi.OUT_EI,
         "poi": _capi.OUT_POI, "pof": _capi.OUT_POF, "ufun": _capi.OUT_UFUN}


def _norm_args(KrigInfo):
    """prediction.py:159-171."""
    if KrigInfo["standardization"] is not True:
        return _capi.NORM_NONE, None, None
    nt = KrigInfo["normtype"]
    if nt == "default":
        return _capi.NORM_DEFAULT, np.asarray(KrigInfo["lb"], dtype=float), np.asarray(KrigInfo["ub"], dtype=float)
    if nt == "std":
        return (_capi.NORM_STD, np.asarray(KrigInfo["X_mean"], dtype=float).ravel(),
                np.asarray(KrigInfo["X_std"], dtype=float).ravel())
    raise ValueError("Kriging model dictionary 'normtype' value: '%s' is not recognised." % nt)


def _check_supported(KrigInfo, num, drm):
    if num is not None:
        raise NotImplementedError("kadal_b200: the multi-objective list layout (num=) is not supported; "
                                  "pass one KrigInfo per objective")
    if drm is not None:
        raise NotImplementedError("kadal_b200: KernelPCA-transformed prediction (drm=) is not supported")
    if KrigInfo.get("norm_y", False) is True:
        raise NotImplementedError("kadal_b200: norm_y=True is broken in the reference (kriging_model.py:120) "
                                  "and not supported")
    idx = np.asarray(KrigInfo["idx"])
    if idx.shape[0] != 1 or np.any(idx != 0):
        raise NotImplementedError("kadal_b200: prediction with TrendOrder >= 1 raises in the reference "
                                  "(prediction.py:250) and is not supported")


def _ensure_predictor(m, KrigInfo):
    """Make the device predictor match the fitted state stored in KrigInfo (Theta, wgkf, U,
    BE, SigmaSqr — prediction.py:148-154), uploading U only when it changed."""
    tok = _pred_token(KrigInfo)
    if m.pred_token != tok:
        s2 = float(np.asarray(KrigInfo["SigmaSqr"]).ravel()[0])
        m.predictor_load(KrigInfo["Theta"], KrigInfo["wgkf"], np.asarray(KrigInfo["U"]), KrigInfo["BE"], s2)
        m.pred_token = tok
        m._pred_keep = (KrigInfo["U"], KrigInfo["BE"])


def predict_arrays(x, KrigInfo, outputs=("pred",), out=None):
    """Bulk predictor: returns ``{name: (npred,) array}`` for names in
    pred | ssqr | s | ei | poi | pof | ufun, plus ``stats``.  ``ei``/``poi`` are returned negated
    exactly like the reference's per-point expressions (batch-level shortcuts are applied by
    :func:`prediction`)."""
    _check_supported(KrigInfo, None, None)
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    want = 0
    for o in outputs:
        want |= _WANT[o]
    normtype, na, nb = _norm_args(KrigInfo)
    limit, limit_ge = 0.0, True
    if want & _capi.OUT_POF:
        lt = KrigInfo["limittype"]
        if lt in (">", ">="):
            limit_ge = True
        elif lt in ("<", "<="):
            limit_ge = False
        else:
            raise ValueError("Limit Type is not available yet!")
        limit = float(KrigInfo["limit"])
    return m.predict(x, want, normtype, na, nb, ymin=m.ymin, limit=limit, limit_ge=limit_ge, out=out)


def prediction(x, KrigInfo, predtypes, num=None, drm=None, **kwargs):
    """Same contract as the reference ``prediction`` (prediction.py:71)."""
    _check_supported(KrigInfo, num, drm)
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if isinstance(predtypes, str):
        predtypes = [predtypes]
    lowered = [p.lower() for p in predtypes]
    need = set()
    for p, orig in zip(lowered, predtypes):
        if p in ("pred", "fpc"):
            need.add("pred")
        elif p in ("ssqr", "ebe"):
            need.add("ssqr")
        elif p == "s":
            need.add("s")
        elif p == "lcb":
            need.update(("pred", "ssqr"))
        elif p == "ei":
            need.update(("ei", "ssqr"))
        elif p == "poi":
            need.add("poi")
        elif p == "pof":
            need.add("pof")
        else:
            raise NotImplementedError("Specified prediction type: '%s' is not recognised." % orig)
    res = predict_arrays(x, KrigInfo, sorted(need))
    npred = x.shape[0]
    col = lambda a: a.reshape(npred, 1)
    outputs = []
    for p in lowered:
        if p == "pred":
            o = col(res["pred"])
        elif p == "ssqr":
            o = col(res["ssqr"])
        elif p == "s":
            o = col(res["s"])
        elif p == "fpc":
            o = np.dot(np.ones((npred, 1)), np.asarray(KrigInfo["BE"]).reshape(1, 1))
        elif p == "lcb":   # (npred,1) - (1,npred) broadcast, as written in prediction.py:268
            o = col(res["pred"]) - np.dot(KrigInfo["sigmalcb"], res["ssqr"].reshape(1, npred))
        elif p == "ebe":   # untransposed in the reference (prediction.py:270)
            o = -res["ssqr"].reshape(1, npred)
        elif p == "ei":
            nzero, anyterm = res["stats"]
            if nzero > 0:                      # `SSqr.all() == 0` (prediction.py:273-274)
                o = 0
            elif not anyterm:                  # random tiny penalty (prediction.py:285-290)
                tiny = np.finfo(float).tiny
                o = -np.array([[np.random.uniform(tiny, tiny * 100)]])
            else:
                o = col(res["ei"])
        elif p == "poi":
            o = col(res["poi"])
        elif p == "pof":
            o = col(res["pof"])
        outputs.append(o)
    if len(outputs) == 1:
        try:
            return outputs[0].item()
        except (ValueError, AttributeError):
            return outputs[0]
    return outputs
