"""Host-side mirror of ``kadal/surrogate_models/supports/trendfunction.py``.

Same names, argument meaning and outputs as the reference (``polytruncation`` :7,
``compute_regression_mat`` :72); the implementation is independent: multi-indices are
enumerated directly in the reference's graded / descending-lexicographic order and the
orthonormal Legendre / Hermite bases use the three-term recurrences instead of the
reference's Gram-Schmidt on coefficient arrays.  O(n p) work that stays on the host (the
reference does the same); for TrendOrder = 0 (ordinary Kriging, every demo) ``F`` is a
column of ones.
"""
import math

import numpy as np


def _compositions_desc(total, parts):
    """All ``parts``-tuples of non-negative ints summing to ``total``, in descending
    lexicographic order (the order of the reference's MonCof :44-69 after the flip)."""
    if parts == 1:
        yield (total,)
        return
    for first in range(total, -1, -1):
        for rest in _compositions_desc(total - first, parts - 1):
            yield (first,) + rest


def polytruncation(nix, nvar, q):
    """Polynomial multi-indices of total order <= nix (q = 1) or hyperbolically truncated
    (q < 1: (sum idx**q)**(1/q) <= nix), one row per basis function (trendfunction.py:7-41)."""
    rows = []
    for deg in range(0, int(nix) + 1):
        rows.extend(_compositions_desc(deg, int(nvar)))
    idx = np.array(rows, dtype=float).reshape(-1, int(nvar))
    if q < 1:
        power = float("inf") if q == 0 else 1.0 / q
        norm = np.sum(idx ** q, 1) ** power
        idx = idx[norm <= (nix + 0.000001), :]
    return idx


def _legendre_all(kmax, t):
    """P_0..P_kmax at t (Bonnet recurrence)."""
    out = [np.ones_like(t)]
    if kmax >= 1:
        out.append(t.copy())
    for k in range(1, kmax):
        out.append(((2 * k + 1) * t * out[k] - k * out[k - 1]) / (k + 1))
    return out


def _hermite_prob_all(kmax, t):
    """Probabilists' He_0..He_kmax at t (He_{k+1} = t He_k - k He_{k-1}; the reference's
    hermite_rec with id = 1, trendfunction.py:252-270)."""
    out = [np.ones_like(t)]
    if kmax >= 1:
        out.append(t.copy())
    for k in range(1, kmax):
        out.append(t * out[k] - k * out[k - 1])
    return out


def compute_regression_mat(idx, X, bounds, polytype):
    """Regression matrix F (nsamp x nbasis) of orthonormal polynomial-chaos bases
    (trendfunction.py:72-111).  polytype[k] = 1: Legendre on [bounds[0,k], bounds[1,k]]
    scaled by sqrt(2 order + 1); = 2: Hermite with the reference's scaling
    2**(-order/2) He_order((x - b0)/sqrt(2 b1**2)) / sqrt(order!)."""
    X = np.asarray(X, dtype=float)
    idx = np.atleast_2d(np.asarray(idx))
    bounds = np.asarray(bounds, dtype=float)
    nsamp, nvar = X.shape
    F = np.ones((nsamp, idx.shape[0]))
    kmax = int(idx.max()) if idx.size else 0
    basis = []
    for k in range(nvar):
        if polytype[k] == 1:
            t = 2 * ((X[:, k] - bounds[0, k]) / (bounds[1, k] - bounds[0, k])) - 1
            vals = _legendre_all(kmax, t)
            basis.append([v * math.sqrt(2 * o + 1) for o, v in enumerate(vals)])
        elif polytype[k] == 2:
            t = (X[:, k] - bounds[0, k]) / np.sqrt(2 * bounds[1, k] ** 2)
            vals = _hermite_prob_all(kmax, t)
            basis.append([v / (2 ** (o / 2) * math.sqrt(math.factorial(o))) for o, v in enumerate(vals)])
        else:
            raise ValueError("polytype must be 1 (Legendre) or 2 (Hermite)")
    for i in range(idx.shape[0]):
        col = np.ones(nsamp)
        for k in range(nvar):
            col = col * basis[k][int(idx[i, k])]
        F[:, i] = col
    return F
