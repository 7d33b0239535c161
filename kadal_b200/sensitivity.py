"""Drop-in for ``kadal/sensitivity_analysis/sobol_ind.py`` on a Kriging surrogate (SURVEY.md §8 row
f3, BASELINE config 4).

``SobolIndices`` keeps the reference's constructor / ``analyze`` contract (sobol_ind.py:6-99) for the
``krigobj`` case.  The reference predicts A, B and every C_i in 10 000-row slices and reduces (N,1)
host vectors; here ``kb200_sobol_sums`` streams the sample through the GPU once per column set and
returns only the Saltelli sums, from which the indices are formed with the reference's formulas
(:84-85, :136-139, :180-182).  Under ``torch.distributed`` every rank passes its row shard
(``parallel.shard_bounds`` order) and the sums are all-reduced.
"""
import ctypes

import numpy as np

from ._capi import as_f64, check, ptr
from .model import model_for
from .prediction import _check_supported, _ensure_predictor, _norm_args


def sobol_sums(A, B, KrigInfo, column_sets, group=None):
    """(base[4], ya_yc[nsets], yb_yc[nsets]) — see include/kb200.h:kb200_sobol_sums.  ``column_sets`` is
    a list of column-index tuples taken from A."""
    _check_supported(KrigInfo, None, None)
    A = as_f64(np.atleast_2d(A))
    B = as_f64(np.atleast_2d(B))
    if A.shape != B.shape:
        raise ValueError("A and B must have the same shape")
    m = model_for(KrigInfo)
    if A.shape[1] != m.d:
        raise ValueError("sample has %d columns, model has %d" % (A.shape[1], m.d))
    _ensure_predictor(m, KrigInfo)
    normtype, na, nb = _norm_args(KrigInfo)
    nsets = len(column_sets)
    masks = np.zeros((max(nsets, 1), m.d), dtype=np.uint8)
    for q, cols in enumerate(column_sets):
        masks[q, list(cols)] = 1
    base, yayc, ybyc = np.zeros(4), np.zeros(max(nsets, 1)), np.zeros(max(nsets, 1))
    na = as_f64(na) if na is not None else None
    nb = as_f64(nb) if nb is not None else None
    if A.shape[0] > 0:
        check(m.lib.kb200_sobol_sums(m.handle, ptr(A), ptr(B), A.shape[0], int(normtype), ptr(na), ptr(nb),
                                     masks.ctypes.data_as(ctypes.c_void_p), nsets, ptr(base), ptr(yayc), ptr(ybyc)))
    from . import parallel
    world, _ = parallel._world(group)
    if world > 1:
        import torch
        import torch.distributed as dist
        dev = parallel.comm_device(group)
        t = torch.from_numpy(np.concatenate([base, yayc, ybyc])).to(dev)
        dist.all_reduce(t, group=group)
        v = t.cpu().numpy()
        base, yayc, ybyc = v[:4], v[4:4 + len(yayc)], v[4 + len(yayc):]
    return base, yayc[:nsets], ybyc[:nsets]


class SobolIndices:
    """sobol_ind.py:6-99 for a Kriging surrogate.  ``A`` / ``B`` may be given (row shards under data
    parallelism, with ``nMC`` the global row count); otherwise they are drawn like the reference does
    (:24-29, a 2*nvar-column Sobol plan split in two)."""

    def __init__(self, nvar, krigobj=None, problem=None, ub=None, lb=None, nMC=2e5, A=None, B=None, group=None):
        if krigobj is None:
            raise NotImplementedError("kadal_b200: SobolIndices needs a Kriging surrogate (krigobj); evaluating an "
                                      "analytical `problem` is not part of the GPU path")
        self.nvar = nvar
        self.krigobj = krigobj
        self.problem = problem
        self.n = int(nMC)
        self.group = group
        if A is None or B is None:
            from .sampling import sampling
            if ub is not None and lb is not None:
                _, init_mat = sampling("sobolnew", self.nvar * 2, self.n, result="real",
                                       upbound=ub, lobound=lb)      # 2*nvar-long bounds, as in the reference
            else:
                init_mat, _ = sampling("sobolnew", self.nvar * 2, self.n)
            A, B = init_mat[:, :self.nvar], init_mat[:, self.nvar:]
        self.A = np.ascontiguousarray(A, dtype=float)
        self.B = np.ascontiguousarray(B, dtype=float)
        self.fo_2 = None
        self.denom = None

    def _info(self):
        return self.krigobj.KrigInfo if hasattr(self.krigobj, "KrigInfo") else self.krigobj

    def analyze(self, first=True, total=False, second=False):
        sets = []
        if first or total or second:
            sets += [(i,) for i in range(self.nvar)]
        pairs = [(i, j) for i in range(self.nvar - 1) for j in range(i + 1, self.nvar)] if second else []
        base, yayc, ybyc = sobol_sums(self.A, self.B, self._info(), sets + pairs, self.group)
        n = self.n
        self.fo_2 = (base[0] / n) ** 2                       # sobol_ind.py:84
        self.denom = (base[1] / n) - self.fo_2               # :85
        indices = dict()
        s1 = np.zeros(self.nvar)
        st = np.zeros(self.nvar)
        if sets:
            for ii in range(self.nvar):
                if first:
                    s1[ii] = ((1 / n) * yayc[ii] - self.fo_2) / self.denom                # :136-137
                if total:
                    st[ii] = 1 - (((1 / n) * ybyc[ii] - self.fo_2) / self.denom)          # :138-139
        if first is True or total is True:
            indices["first"], indices["total"] = s1, st
        if second is True:
            s1 = indices["first"]          # KeyError when neither first nor total was asked, as in the reference (:90-91)
            s2 = dict()
            for q, (ii, jj) in enumerate(pairs):
                vij = (1 / n) * yayc[self.nvar + q] - self.fo_2                           # :180
                s2["x" + str(ii + 1) + "-x" + str(jj + 1)] = (vij / self.denom) - s1[ii] - s1[jj]
            indices["second"] = s2
        return indices
