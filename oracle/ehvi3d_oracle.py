"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's 3-objective EHVI (KMAC, "multi
slice-update"): kadal/extern/HDYE_3D_Update/ehvi_multi.cc:113-341 as reached from
optim_tools/ehvi/EHVIcomputation.py:235 through kmac_wrapper.cpp:73-118.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Pinned: against the reference's own known answers (test_single.py:21, test_multi.py:26-27) and against the
reference's C++ itself, compiled from /root/reference by oracle/kmac_ref/Makefile into
oracle/_ref/libkmac_ref.so (tests/test_ehvi3d.py; goldens tests/golden/ehvi3d.npz).

What is restated, including the parts one would not design:
  * maximisation convention, infinite upper cells, reference point = lower corner;
  * the constants are the reference's truncated ones (ehvi_consts.h:9-10: sqrt(2) and 1/sqrt(2 pi) to
    10-11 digits), so results differ from the mathematically exact EHVI in the 10th digit -- and agree with
    the reference to rounding;
  * ehvi_multi.cc:296-308 leaves the candidate loop with ``break`` (not ``continue``) as soon as ONE
    candidate's partial improvement in the cell is below 1e-15: every later candidate of the batch loses that
    cell.  The result of a candidate therefore depends on the candidates in front of it; the reference's own
    test_multi.py pins this (its third answer, 0, is 3.974 when that candidate is evaluated alone).
    ``independent=True`` evaluates every candidate as if it were alone instead;
  * contributions are clipped at zero cell by cell (:330), products below 1e-15 are dropped (:311).
"""
import numpy as np
from scipy.special import erf

SQRT_TWO = 1.41421356237          # ehvi_consts.h:9
SQRT_TWOPI_NEG = 0.3989422804     # ehvi_consts.h:10
THRESH = 0.000000000000001        # ehvi_multi.cc:298,303,308,311


def gausspdf(x):                   # helper.h:76-78
    return SQRT_TWOPI_NEG * np.exp(-(x * x) / 2)


def gausscdf(x):                   # helper.h:81-83
    return 0.5 * (1 + erf(x / SQRT_TWO))


def exipsi(fmax, corner, mu, s):   # helper.h:86-88
    z = (corner - mu) / s
    return (s * gausspdf(z)) + ((fmax - mu) * gausscdf(z))


def kmac_front(y_par):
    """kmac_wrapper.cpp:24-34, 80-91: points are fed one by one; a new point removes the earlier points it
    dominates or equals (it is itself never rejected)."""
    front = []
    for p in np.asarray(y_par, dtype=float).reshape(-1, 3):
        front = [q for q in front if not (p[0] >= q[0] and p[1] >= q[1] and p[2] >= q[2])]
        front.append(p)
    return np.array(front).reshape(-1, 3)


def front_tables(P, r):
    """Geometry that depends on the front only (ehvi_multi.cc:139-196).  Returns sorted coordinates X, Y, Z,
    hd[x, y] (z-rank of the highest point dominating grid node (x, y), -1 if none), xlim[y, z], ylim[x, z]."""
    n = len(P)
    ox, oy, oz = (np.argsort(P[:, c], kind="stable") for c in range(3))
    xr, yr, zr = (np.empty(n, dtype=int) for _ in range(3))
    xr[ox], yr[oy], zr[oz] = np.arange(n), np.arange(n), np.arange(n)
    X, Y, Z = P[ox, 0], P[oy, 1], P[oz, 2]
    hd = -np.ones((n, n), dtype=int)
    xlim, ylim = np.zeros((n, n)), np.zeros((n, n))
    for x in range(n):
        for y in range(n):
            m = (xr >= x) & (yr >= y)
            if m.any():
                hd[x, y] = zr[m].max()
    for a in range(n):
        for z in range(n):
            m = (yr >= a) & (zr >= z)
            if m.any():
                xlim[a, z] = P[m, 0].max() - r[0]
            m = (xr >= a) & (zr >= z)
            if m.any():
                ylim[a, z] = P[m, 1].max() - r[1]
    return X, Y, Z, hd, xlim, ylim


def ehvi3d_multi(y_par, ref_point, mean, sdev, independent=False, prefilter=True):
    """kmac.ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev) -> [n_pop]."""
    P = kmac_front(y_par) if prefilter else np.asarray(y_par, dtype=float).reshape(-1, 3)
    r = np.asarray(ref_point, dtype=float).reshape(3)
    MU = np.asarray(mean, dtype=float).reshape(-1, 3)
    S = np.asarray(sdev, dtype=float).reshape(-1, 3)
    n, npop = len(P), len(MU)
    X, Y, Z, hd, xlim, ylim = front_tables(P, r)
    lo = [np.r_[r[c], (X, Y, Z)[c]] for c in range(3)]          # cl of cell 0..n
    hi = [np.r_[(X, Y, Z)[c], np.inf] for c in range(3)]        # cu of cell 0..n
    with np.errstate(all="ignore"):
        # per axis, cell and candidate: partial improvement, probability mass, distance term (:294-320)
        E, G, EX = [], [], []
        for c in range(3):
            cl, cu = lo[c][:, None], hi[c][:, None]
            mu, s = MU[None, :, c], S[None, :, c]
            e = exipsi(r[c], cl, mu, s) - exipsi(r[c], cu, mu, s)
            g = gausscdf((cu - mu) / s) - gausscdf((cl - mu) / s)
            E.append(e), G.append(g), EX.append(e - (g * (cl - r[c])))
        idx = np.arange(npop)
        if independent:
            first = [np.full(n + 1, npop) for _ in range(3)]
        else:   # first candidate that leaves the loop, per axis and cell
            first = [np.where((E[c] < THRESH).any(axis=1), (E[c] < THRESH).argmax(axis=1), npop) for c in range(3)]
        answer = np.zeros(npop)
        slice_ = np.zeros((n, n))
        chunk = np.zeros((n, n))
        for z in range(n + 1):
            if z > 0:
                chunk = chunk + slice_ * (hi[2][z - 1] - lo[2][z - 1])               # :212-215
            new = np.zeros((n, n))
            for y in range(n):                                                       # :217-236
                for x in range(n):
                    if hd[x, y] < z:
                        if x > 0 and y > 0:
                            new[x, y] = (new[x, y - 1] - new[x - 1, y - 1]) + new[x - 1, y]
                        elif y > 0:
                            new[x, y] = new[x, y - 1]
                        elif x > 0:
                            new[x, y] = new[x - 1, y]
                    else:
                        new[x, y] = (X[x] - r[0]) * (Y[y] - r[1])
            slice_ = new
            len2 = hi[2][z] - lo[2][z]
            for y in range(n + 1):
                len1 = hi[1][y] - lo[1][y]
                for x in range(n + 1):
                    len0 = hi[0][x] - lo[0][x]
                    if len0 == 0 or len1 == 0 or len2 == 0 or (x < n and y < n and hd[x, y] >= z):
                        continue                                                     # :253-254
                    if x > 0 and y > 0:                                              # :258-269
                        sminus = chunk[x - 1, y - 1]
                        sl0 = 0.0 if x == n else (chunk[x, y - 1] - sminus) / len0
                        sl1 = 0.0 if y == n else (chunk[x - 1, y] - sminus) / len1
                        sl2 = slice_[x - 1, y - 1]
                    else:
                        sminus = 0.0
                        sl0 = 0.0 if (y == 0 or x == n) else (chunk[x, y - 1] - sminus) / len0
                        sl1 = 0.0 if (x == 0 or y == n) else (chunk[x - 1, y] - sminus) / len1
                        sl2 = 0.0
                    v0 = 0.0 if (y == n or z == n) else xlim[y, z]                   # :271-282
                    v1 = 0.0 if (x == n or z == n) else ylim[x, z]
                    v2 = 0.0 if (x == n or y == n) else (0.0 if hd[x, y] == -1 else Z[hd[x, y]] - r[2])
                    e1, e2, e3 = E[0][x], E[1][y], E[2][z]
                    g1, g2, g3 = G[0][x], G[1][y], G[2][z]
                    ex1, ex2, ex3 = EX[0][x], EX[1][y], EX[2][z]
                    live = idx < min(first[0][x], first[1][y], first[2][z])
                    if independent:
                        live = ~((e1 < THRESH) | (e2 < THRESH) | (e3 < THRESH))
                    psiprod = e1 * e2 * e3
                    tot = (psiprod - (sminus * g1 * g2 * g3)
                           - (sl0 * g2 * g3 * ex1) - v0 * ex2 * ex3 * g1
                           - (sl1 * g1 * g3 * ex2) - v1 * ex1 * ex3 * g2
                           - (sl2 * g1 * g2 * ex3) - v2 * ex1 * ex2 * g3)            # :321-325
                    add = live & (psiprod > THRESH) & (tot > 0)
                    answer = np.where(add, answer + tot, answer)
    return answer
