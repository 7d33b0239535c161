"""CPU restatement of KADAL's 2-objective expected hyper-volume improvement.

TEST INFRASTRUCTURE ONLY (parity checker); the product never imports it.

====================  =====================================================================
restated function     reference (file:line, relative to /root/reference)
====================  =====================================================================
``gausspdf``          kadal/optim_tools/ehvi/gaussfcn.py:4-6
``gausscdf``          gaussfcn.py:8-10  (``math.erf``)
``exipsi``            kadal/optim_tools/ehvi/exipsi.py:4-6
``hvolume2d``         kadal/optim_tools/ehvi/hvolume2d.py:3-14
``cell_table``        exi2d.py:8-61   (everything of the cell loop that does not depend on mu, s)
``exi2d``             exi2d.py:7-84
``ehvicalc_vec``      kadal/optim_tools/ehvi/EHVIcomputation.py:129-186 (serial branch)
====================  =====================================================================

The reference walks the (k+1)(k+2)/2 staircase cells of the sorted Pareto front for every
candidate; the geometry of a cell (its bounds, the staircase corner fmax and the dominated
hyper-volume ``sPlus``) depends only on the front, so it is tabulated once here and the
candidate loop evaluates the two marginal integrals per cell.  Per cell the floating-point
operations and their order are the reference's.  Pinned by tests/golden/ehvi2d.npz (outputs of
the live reference) and, in the build container, against the live ``exi2d`` itself.

Reference behaviours reproduced on purpose: callers pass the kriging *variance* ``SSqr`` as
``s`` (EHVIcomputation.py:57,176), negative cells are clamped to 0, the final reduction is
``np.sum(np.sum(c, axis=0))``.
"""
from math import erf

import numpy as np

_INV_SQRT_2PI = 1 / np.sqrt(2 * np.pi)
_SQRT2 = np.sqrt(2)


def gausspdf(x):
    return _INV_SQRT_2PI * np.exp(-(x ** 2) / 2)


def gausscdf(x):
    return 0.5 * (1 + erf(x / _SQRT2))


def exipsi(a, b, m, s):
    z = (b - m) / s
    return s * gausspdf(z) + (a - m) * gausscdf(z)


def hvolume2d(P, x):
    """Area dominated by the rows of P inside the box with upper corner x (P sorted on column 0)."""
    if np.size(P, axis=0) == 0:
        return 0
    S = P[np.argsort(P[:, 0]), :]
    h = (x[0] - S[0, 0]) * (x[1] - S[0, 1])
    for i in range(1, len(S)):
        h = h + (x[0] - S[i, 0]) * (S[i - 1, 1] - S[i, 1])
    return h


def sorted_front(P):
    """(S, c1, c2): front sorted on objective 1, and both coordinate lists ascending (exi2d.py:8-13)."""
    P = np.asarray(P, dtype=float)
    S = P[np.argsort(P[:, 0]), :]
    return S, np.sort(S[:, 0]), np.sort(S[:, 1])


def cell_table(P, r):
    """Geometry of the staircase cells: arrays of shape (k+1, k+1) indexed [i+1, j+1] with a
    validity mask (row i+1 has k+1-max(i+1,0) cells... exi2d.py:16-21), holding fmax1, fmax2,
    cL1, cU1, cL2, cU2 and sPlus."""
    S, c1, c2 = sorted_front(P)
    k = S.shape[0]
    shp = (k + 1, k + 1)
    T = {nm: np.zeros(shp) for nm in ("fmax1", "fmax2", "cL1", "cU1", "cL2", "cU2", "sPlus")}
    valid = np.zeros(shp, dtype=bool)
    for i in range(-1, k):
        ii = 0 if i == -1 else i + 1
        for j in range(-1, k - ii):
            fmax2 = r[1] if j == -1 else c2[k - j - 1]
            fmax1 = r[0] if i == -1 else c1[k - i - 1]
            cL1 = -np.inf if j == -1 else c1[j]
            cL2 = -np.inf if i == -1 else c2[i]
            cU1 = r[0] if j == k - 1 else c1[j + 1]
            cU2 = r[1] if i == k - 1 else c2[i + 1]
            dom = (cU1 <= S[:, 0]) & (cU2 <= S[:, 1])
            sPlus = hvolume2d(S[dom][::-1], np.array([fmax1, fmax2])) if dom.any() else 0
            a, b = i + 1, j + 1
            valid[a, b] = True
            for nm, v in (("fmax1", fmax1), ("fmax2", fmax2), ("cL1", cL1), ("cU1", cU1), ("cL2", cL2),
                          ("cU2", cU2), ("sPlus", sPlus)):
                T[nm][a, b] = v
    T["valid"] = valid
    return T


def exi2d(P, r, mu, s, table=None):
    """Expected hyper-volume improvement of N(mu, diag(s)) over the front P w.r.t. reference point r."""
    T = cell_table(P, r) if table is None else table
    k1 = T["valid"].shape[0]
    c = np.zeros((k1, k1))
    for a in range(k1):
        for b in range(k1):
            if not T["valid"][a, b]:
                continue
            f1, f2 = T["fmax1"][a, b], T["fmax2"][a, b]
            Psi1 = exipsi(f1, T["cU1"][a, b], mu[0], s[0]) - exipsi(f1, T["cL1"][a, b], mu[0], s[0])
            Psi2 = exipsi(f2, T["cU2"][a, b], mu[1], s[1]) - exipsi(f2, T["cL2"][a, b], mu[1], s[1])
            G1 = gausscdf((T["cU1"][a, b] - mu[0]) / s[0]) - gausscdf((T["cL1"][a, b] - mu[0]) / s[0])
            G2 = gausscdf((T["cU2"][a, b] - mu[1]) / s[1]) - gausscdf((T["cL2"][a, b] - mu[1]) / s[1])
            v = Psi1 * Psi2 - T["sPlus"][a, b] * G1 * G2
            c[a, b] = 0 if v < 0 else v
    return np.sum(np.sum(c, axis=0))


def exi2d_batch(P, r, MU, S):
    T = cell_table(P, r)
    return np.array([exi2d(P, r, MU[i], S[i], T) for i in range(len(MU))])


def ehvicalc_vec(pred, SSqr, y_par, ref_point):
    """EHVIcomputation.py:169-175 given the [n_pop, 2] predictions: hv = -exi2d(...)."""
    return -1 * exi2d_batch(y_par, ref_point, pred, SSqr)
