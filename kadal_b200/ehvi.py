"""Drop-in for KADAL's 2-objective EHVI scoring (``kadal/optim_tools/ehvi``): SURVEY.md §8 row f1.

``exi2d(P, r, mu, s)`` keeps the reference's signature (exi2d.py:7); ``exi2d_batch`` is the additive
bulk form.  ``ehvicalc`` / ``ehvicalc_vec`` mirror EHVIcomputation.py:33-61 and :129-186: the
objectives' ``pred`` and ``SSqr`` come from the batched GPU predictor (one call per objective for
the whole population instead of n_pop x n_obj Python calls), the expected improvement of every
candidate from ``kb200_ehvi2d``.  Like the reference, the kriging *variance* is handed to exi2d
where a standard deviation is meant (EHVIcomputation.py:57,176) — parity, not a fix.

``ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev)`` stands in for the reference's C++
extension ``kmac`` (kadal/extern/HDYE_3D_Update/kmac_wrapper.cpp:73-118), ``ehvicalc_kmac3d`` mirrors
EHVIcomputation.py:189-251 on top of it (``kb200_ehvi3d``).  ``mode="reference"`` (default) reproduces
KMAC's multi slice-update bit of behaviour for bit of behaviour, including its order-dependent early exit
(include/kb200.h); ``mode="independent"`` scores every candidate as if it were alone.
"""
import numpy as np

from . import _capi
from ._capi import as_f64, check, ptr


def _device():
    import os
    lib = _capi.require_gpu()
    dev = int(os.environ.get("KADAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return lib, dev % max(1, lib.kb200_device_count())


def exi2d_batch(P, r, MU, S, sign=1.0):
    """sign * exi2d(P, r, MU[i], S[i]) for every row i; P is the [k, 2] Pareto front, r the reference point."""
    lib, dev = _device()
    P = as_f64(np.asarray(P, dtype=float).reshape(-1, 2))
    r = as_f64(np.asarray(r, dtype=float).reshape(-1))
    MU = as_f64(np.atleast_2d(MU))
    S = as_f64(np.atleast_2d(S))
    if r.shape[0] != 2 or MU.shape[1] != 2 or S.shape != MU.shape:
        raise ValueError("exi2d: two objectives expected (r (2,), mu (n,2), s (n,2))")
    out = np.empty(MU.shape[0])
    check(lib.kb200_ehvi2d(dev, ptr(P), P.shape[0], ptr(r), ptr(MU), ptr(S), MU.shape[0], float(sign), ptr(out)))
    return out


def exi2d(P, r, mu, s):
    """Expected hyper-volume improvement of one candidate (exi2d.py:7-84)."""
    return exi2d_batch(P, r, np.asarray(mu, dtype=float).reshape(1, 2), np.asarray(s, dtype=float).reshape(1, 2))[0]


_EHVI3D_MODES = {"reference": 0, "independent": 1, 0: 0, 1: 1}


def ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev, mode="reference", sign=1.0):
    """kmac.ehvi3d_sliceupdate_multi (maximisation): EHVI of every row of mean_vector / std_dev [n_pop, 3] over
    the front y_par [n_par, 3] with lower corner ref_point."""
    lib, dev = _device()
    if mode not in _EHVI3D_MODES:
        raise ValueError("ehvi3d mode must be 'reference' or 'independent'")
    P = as_f64(np.asarray(y_par, dtype=float).reshape(-1, 3))
    r = as_f64(np.asarray(ref_point, dtype=float).reshape(-1))
    MU = as_f64(np.atleast_2d(mean_vector))
    S = as_f64(np.atleast_2d(std_dev))
    if r.shape[0] != 3 or MU.shape[1] != 3 or S.shape != MU.shape:
        raise ValueError("ehvi3d: three objectives expected (ref (3,), mean (n,3), std (n,3))")
    out = np.empty(MU.shape[0])
    check(lib.kb200_ehvi3d(dev, ptr(P), P.shape[0], ptr(r), ptr(MU), ptr(S), MU.shape[0], _EHVI3D_MODES[mode],
                           float(sign), ptr(out)))
    return out


def ehvicalc_kmac3d(x, y_par, moboInfo, kriglist, pool=None, mode="reference"):
    """EHVIcomputation.py:189-251: -EHVI (KMAC, 3 objectives) for a population x [n_pop, n_dv]; a float for a
    1-D x.  Like the reference the kriging variance SSqr is handed over where KMAC expects a standard
    deviation (:235), and the inputs are negated because KMAC maximises.  ``pool`` is ignored."""
    x = np.asarray(x, dtype=float)
    reshape = x.ndim == 1
    if reshape:
        x = x.reshape(1, -1)
    if len(kriglist) != 3:
        raise ValueError("ehvicalc_kmac3d needs three objectives")
    pred, SSqr = _pred_ssqr(x, kriglist)
    ref_point = np.asarray(moboInfo["refpoint"], dtype=float)
    hv = ehvi3d_sliceupdate_multi(-np.asarray(y_par, dtype=float), -ref_point, -pred, SSqr, mode=mode, sign=-1.0)
    return hv[0] if reshape else hv


def _pred_ssqr(x, kriglist):
    from .prediction import predict_arrays
    n_pop = x.shape[0]
    pred = np.zeros([n_pop, len(kriglist)])
    SSqr = np.zeros([n_pop, len(kriglist)])
    for j, krig in enumerate(kriglist):
        info = krig.KrigInfo if hasattr(krig, "KrigInfo") else krig
        res = predict_arrays(x, info, ("pred", "ssqr"))
        pred[:, j] = res["pred"]
        SSqr[:, j] = res["ssqr"]
    return pred, SSqr


def ehvicalc_vec(x, y_par, moboInfo, kriglist, pool=None):
    """EHVIcomputation.py:129-186: -EHVI for a population x [n_pop, n_dv] (a float for a 1-D x).
    ``pool`` is accepted and ignored (CUDA contexts do not survive fork)."""
    x = np.asarray(x, dtype=float)
    reshape = x.ndim == 1
    if reshape:
        x = x.reshape(1, -1)
    if len(kriglist) != 2:
        raise NotImplementedError("kadal_b200: ehvicalc_vec is the 2-objective exi2d path (EHVIcomputation.py:129); "
                                  "use ehvicalc_kmac3d for three objectives")
    pred, SSqr = _pred_ssqr(x, kriglist)
    hv = exi2d_batch(y_par, moboInfo["refpoint"], pred, SSqr, sign=-1.0)
    return hv[0] if reshape else hv


def ehvicalc(x, ypar, moboInfo, kriglist):
    """EHVIcomputation.py:16-61 (single candidate, with the reference's random tiny penalty at 0)."""
    HV = ehvicalc_vec(np.asarray(x, dtype=float).reshape(-1), ypar, moboInfo, kriglist)
    if HV == 0:
        tiny = np.finfo("float").tiny
        HV = np.random.uniform(tiny, tiny * 100)
    return HV
