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


def _masks(column_sets, d):
    nsets = len(column_sets)
    masks = np.zeros((max(nsets, 1), d), dtype=np.uint8)
    for q, cols in enumerate(column_sets):
        masks[q, list(cols)] = 1
    return masks


def _all_reduce_sums(base, yayc, ybyc, group):
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
    return base, yayc, ybyc


def sobol_sums(A, B, KrigInfo, column_sets, group=None, stream=0):
    """(base[4], ya_yc[nsets], yb_yc[nsets]) — see include/kb200.h:kb200_sobol_sums.  ``column_sets`` is
    a list of column-index tuples taken from A.  ``A`` / ``B`` are host arrays, or raw device addresses (ints) with
    ``A = (ptr, N)``-style tuples ``(data_ptr, nrows)`` for samples that are already resident in HBM
    (``kb200_sobol_sums_dev``)."""
    _check_supported(KrigInfo, None, None)
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    normtype, na, nb = _norm_args(KrigInfo)
    nsets = len(column_sets)
    masks = _masks(column_sets, m.d)
    base, yayc, ybyc = np.zeros(4), np.zeros(max(nsets, 1)), np.zeros(max(nsets, 1))
    na = as_f64(na) if na is not None else None
    nb = as_f64(nb) if nb is not None else None
    if isinstance(A, tuple):
        (pa, N), (pb, Nb) = A, B
        if N != Nb:
            raise ValueError("A and B must have the same number of rows")
        if N > 0:
            check(m.lib.kb200_sobol_sums_dev(m.handle, ctypes.c_void_p(pa), ctypes.c_void_p(pb), int(N), int(normtype), ptr(na),
                                             ptr(nb), masks.ctypes.data_as(ctypes.c_void_p), nsets, ptr(base), ptr(yayc),
                                             ptr(ybyc), ctypes.c_void_p(stream)))
    else:
        A = as_f64(np.atleast_2d(A))
        B = as_f64(np.atleast_2d(B))
        if A.shape != B.shape:
            raise ValueError("A and B must have the same shape")
        if A.shape[1] != m.d:
            raise ValueError("sample has %d columns, model has %d" % (A.shape[1], m.d))
        if A.shape[0] > 0:
            check(m.lib.kb200_sobol_sums(m.handle, ptr(A), ptr(B), A.shape[0], int(normtype), ptr(na), ptr(nb),
                                         masks.ctypes.data_as(ctypes.c_void_p), nsets, ptr(base), ptr(yayc), ptr(ybyc)))
    base, yayc, ybyc = _all_reduce_sums(base, yayc, ybyc, group)
    return base, yayc[:nsets], ybyc[:nsets]


def sobol_sums_generated(nMC, KrigInfo, column_sets, lb=None, ub=None, group=None):
    """The same sums with A and B drawn ON THE DEVICE exactly as the reference draws them (sobol_ind.py:24-29: one Sobol
    plan of 2*nvar columns, row 0 dropped, A | B its halves, realval-scaled to [lb, ub] of length 2*nvar): bit-identical to
    ``sampling('sobolnew', 2*nvar, nMC, ...)`` on the host, nothing but the sums crosses PCIe
    (``kb200_sobol_sums_gen``).  Under ``torch.distributed`` every rank generates its own contiguous row shard
    (``parallel.shard_bounds``) and the sums are all-reduced.  Returns None when the host generator in use cannot be
    reproduced on the device (the ``sobolsampling`` package is installed): generate on the host then."""
    from . import parallel
    from .sampling import sobol_direction_numbers
    _check_supported(KrigInfo, None, None)
    m = model_for(KrigInfo)
    dn = sobol_direction_numbers(2 * m.d, int(nMC) + 1)
    if dn is None:
        return None
    v, bits = dn
    _ensure_predictor(m, KrigInfo)
    normtype, na, nb = _norm_args(KrigInfo)
    nsets = len(column_sets)
    masks = _masks(column_sets, m.d)
    base, yayc, ybyc = np.zeros(4), np.zeros(max(nsets, 1)), np.zeros(max(nsets, 1))
    na = as_f64(na) if na is not None else None
    nb = as_f64(nb) if nb is not None else None
    lo = as_f64(lb) if lb is not None else None
    hi = as_f64(ub) if ub is not None else None
    if (lo is None) != (hi is None) or (lo is not None and (lo.shape != (2 * m.d,) or hi.shape != (2 * m.d,))):
        raise ValueError("lb and ub must both have 2*nvar entries (the A | B halves of one plan, as in the reference)")
    world, rank = parallel._world(group)
    r0, r1 = parallel.shard_bounds(int(nMC), world, rank)
    if r1 > r0:
        check(m.lib.kb200_sobol_sums_gen(m.handle, v.ctypes.data_as(ctypes.c_void_p), bits, 1 + r0, r1 - r0, ptr(lo), ptr(hi),
                                         int(normtype), ptr(na), ptr(nb), masks.ctypes.data_as(ctypes.c_void_p), nsets,
                                         ptr(base), ptr(yayc), ptr(ybyc)))
    base, yayc, ybyc = _all_reduce_sums(base, yayc, ybyc, group)
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
        self.ub, self.lb = ub, lb
        self._A = None if A is None else np.ascontiguousarray(A, dtype=float)
        self._B = None if B is None else np.ascontiguousarray(B, dtype=float)
        self.fo_2 = None
        self.denom = None

    def _draw(self):
        """The reference's sample matrices on the host (sobol_ind.py:24-29); only materialised when somebody asks for
        ``.A`` / ``.B`` or the device generator cannot reproduce the host one."""
        from .sampling import sampling
        if self.ub is not None and self.lb is not None:
            _, init_mat = sampling("sobolnew", self.nvar * 2, self.n, result="real",
                                   upbound=self.ub, lobound=self.lb)      # 2*nvar-long bounds, as in the reference
        else:
            init_mat, _ = sampling("sobolnew", self.nvar * 2, self.n)
        self._A = np.ascontiguousarray(init_mat[:, :self.nvar], dtype=float)
        self._B = np.ascontiguousarray(init_mat[:, self.nvar:], dtype=float)

    @property
    def A(self):
        if self._A is None:
            self._draw()
        return self._A

    @property
    def B(self):
        if self._B is None:
            self._draw()
        return self._B

    def _info(self):
        return self.krigobj.KrigInfo if hasattr(self.krigobj, "KrigInfo") else self.krigobj

    def analyze(self, first=True, total=False, second=False):
        sets = []
        if first or total or second:
            sets += [(i,) for i in range(self.nvar)]
        pairs = [(i, j) for i in range(self.nvar - 1) for j in range(i + 1, self.nvar)] if second else []
        res = None
        if self._A is None and self._B is None:      # no sample given: draw it on the device, as the reference would on the host
            res = sobol_sums_generated(self.n, self._info(), sets + pairs, self.lb, self.ub, self.group)
        if res is None:
            res = sobol_sums(self.A, self.B, self._info(), sets + pairs, self.group)
        base, yayc, ybyc = res
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
