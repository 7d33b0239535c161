"""Drop-in for ``kadal/surrogate_models/supports/likelihood_func.py``.

``likelihood(x, KrigInfo, mode, trainvar)`` keeps the reference's signature, return values,
KrigInfo side effects and error types (likelihood_func.py:6-118); ``likelihood_batch`` is the
additive population entry point that the optimisers in :mod:`kadal_b200.kriging_model` use.
All arithmetic runs in the CUDA library (``kb200_lik_batch`` / ``kb200_fit_state``).
"""
import numpy as np

from .model import model_for, append_source


def _fixed_nugget(KrigInfo):
    """KrigInfo['nugget'] as the fixed log10 nugget (likelihood_func.py:124); when the nugget
    is tuned the value comes from x and this one is ignored."""
    nug = KrigInfo["nugget"]
    if isinstance(nug, (list, tuple, np.ndarray)):
        return float(np.asarray(nug, dtype=float).ravel()[0])
    return float(nug)


def _as_vector(x, KrigInfo):
    """likelihood_func.py:60-64: scalars are wrapped; iso_gaussian takes a single theta."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 0:
        x = x.reshape(1)
    return x


def likelihood_batch(Xpop, KrigInfo, trainvar=False):
    """Negative ln-likelihood of every row of ``Xpop`` (P x nbhyp) in one GPU call.

    Returns ``(nll, info)``: ``nll[i]`` equals ``likelihood(Xpop[i], KrigInfo, 'default',
    trainvar)`` of the reference; candidates whose correlation matrix is not positive
    definite get ``+inf`` and ``info[i] != 0`` (the reference raises ``LinAlgError`` there,
    which would kill a whole population).  KrigInfo is not modified.
    """
    Xpop = np.atleast_2d(np.asarray(Xpop, dtype=float))
    m = model_for(KrigInfo)
    return m.lik_batch(Xpop, trainvar is True, _fixed_nugget(KrigInfo))


def _layout(m, nbhyp, tv):
    """nuggetset (likelihood_func.py:121-165): which optional entries x carries.
    Returns (offset of the first optional entry, tuned nugget?, kernel weights?)."""
    base = (1 if m.iso else m.ntheta) + (1 if tv else 0)
    extra = nbhyp - base
    if extra == 0:
        return base, False, False
    if extra == 1:
        return base, True, False
    if extra == m.nkernel:
        return base, False, True
    if extra == m.nkernel + 1:
        return base, True, True
    raise ValueError("Nugget setting is not available")


def _side_effects(KrigInfo, m, xv, tv):
    """KrigInfo['Theta'|'nugget'|'wgkf'] are rewritten on every call (likelihood_func.py:82-84)."""
    base, tuned, comp = _layout(m, len(xv), tv)
    KrigInfo["Theta"] = np.array([xv[0]] * m.ntheta) if m.iso else xv[0:m.ntheta]
    if tuned:
        KrigInfo["nugget"] = xv[base]
    if comp:
        w = xv[base + (1 if tuned else 0):base + (1 if tuned else 0) + m.nkernel]
        KrigInfo["wgkf"] = w / np.sum(w)
    else:
        KrigInfo["wgkf"] = np.array([1])


def likelihood(x, KrigInfo, mode="default", trainvar=True):
    """Same contract as the reference ``likelihood`` (likelihood_func.py:6).

    mode='default' returns NegLnLike; mode='all' returns KrigInfo with U, Psi, BE, SigmaSqr,
    NegLnLike filled in (:107-111).  Theta / nugget / wgkf are written on every call (:82-84).
    Deviation: in 'default' mode U/Psi/BE/SigmaSqr are *not* refreshed (that would cost two
    n x n device->host copies per evaluation); every caller in the reference reads them only
    after a mode='all' call (kriging_model.py:258, MOBO.py:353, krigloocv2.py:25-27).
    A non positive definite correlation matrix raises ``numpy.linalg.LinAlgError`` like the
    reference's ``np.linalg.cholesky`` (:175); the blocking ``input()`` of :211 is not kept.
    """
    lmode = mode.lower()
    if lmode not in ("default", "all"):
        raise TypeError("Only have two modes, default and all, default return NegLnLike, all return KrigInfo")
    m = model_for(KrigInfo)
    xv = _as_vector(x, KrigInfo)
    nug_fixed = _fixed_nugget(KrigInfo)
    tv = trainvar is True
    _side_effects(KrigInfo, m, xv, tv)
    if lmode == "default":
        nll, info = m.lik_batch(xv[None, :], tv, nug_fixed)
        if info[0] != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        KrigInfo["NegLnLike"] = nll[0]
        return nll[0]
    # one more sample under the same hyper-parameters (believer / sequential updates): extend the cached factor
    st = m.fit_state(xv, tv, nug_fixed, append_to=append_source(m, xv, tv, nug_fixed))
    if st["info"] != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    KrigInfo["U"] = st["U"]
    KrigInfo["Psi"] = st["Psi"]
    KrigInfo["BE"] = st["BE"].reshape(-1, 1)
    KrigInfo["SigmaSqr"] = np.float64(st["sigma2"]) if tv else np.array([[st["sigma2"]]])
    KrigInfo["NegLnLike"] = np.float64(st["nll"])
    # the predictor installed by kb200_fit_state belongs to exactly these arrays
    if m.p == 1:
        m.pred_token = _pred_token(KrigInfo)
        m._pred_keep = (KrigInfo["U"], KrigInfo["BE"])
    return KrigInfo


def _pred_token(KrigInfo):
    """Identity of the fitted state prediction() reads (prediction.py:148-154)."""
    s2 = KrigInfo["SigmaSqr"]
    return (id(KrigInfo["U"]), np.asarray(KrigInfo["Theta"], dtype=float).tobytes(),
            np.asarray(KrigInfo["BE"], dtype=float).tobytes(), float(np.asarray(s2).ravel()[0]),
            np.asarray(KrigInfo["wgkf"], dtype=float).tobytes())
