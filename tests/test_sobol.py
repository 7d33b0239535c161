"""Sobol indices through a Kriging surrogate (BASELINE config 4).

GPU: ``kadal_b200.sensitivity.SobolIndices`` (device-side Saltelli sums) against the reference's own
estimator restated on the oracle's predictions (sobol_ind.py:84-85, 136-139, 180-182; np.sum over
the full prediction vectors).  The sums are accumulated in a different order than numpy's pairwise
sum; the indices are compared to 1e-9 absolute (they are O(1) ratios of O(N) sums)."""
import copy

import numpy as np
import pytest

from oracle import kriging_oracle as ko

pytestmark = pytest.mark.gpu


def reference_indices(A, B, info, second):
    """sobol_ind.py:45-141 with the oracle as the surrogate."""
    n, d = A.shape
    pred = lambda x: ko.prediction(x, info, ["pred"])
    ya, yb = pred(A), pred(B)
    fo_2 = (np.sum(ya) / n) ** 2
    denom = (np.sum(ya ** 2) / n) - fo_2
    s1, st = np.zeros(d), np.zeros(d)
    for ii in range(d):
        C = copy.deepcopy(B)
        C[:, ii] = A[:, ii]
        yci = pred(C)
        s1[ii] = ((1 / n) * np.sum(ya * yci) - fo_2) / denom
        st[ii] = 1 - (((1 / n) * np.sum(yb * yci) - fo_2) / denom)
    s2 = {}
    if second:
        for ii in range(d - 1):
            for jj in range(ii + 1, d):
                C = copy.deepcopy(B)
                C[:, ii] = A[:, ii]
                C[:, jj] = A[:, jj]
                vij = (1 / n) * np.sum(ya * pred(C)) - fo_2
                s2["x%d-x%d" % (ii + 1, jj + 1)] = (vij / denom) - s1[ii] - s1[jj]
    return s1, st, s2


def test_sobol_indices_match_the_reference_estimator():
    import kadal_b200
    from kadal_b200.sensitivity import SobolIndices, sobol_sums
    rng = np.random.default_rng(17)
    n, d = 150, 5
    X = rng.uniform(-1, 1, (n, d))
    y = (np.sin(2 * X[:, 0]) + 0.7 * X[:, 1] ** 2 + 0.3 * X[:, 0] * X[:, 2])[:, None] + 3.0
    info = ko.make_kriginfo(X, y)
    kadal_b200.likelihood(np.full(d, 0.15), info, "all", False)
    ref = copy.deepcopy(info)
    ko.likelihood(np.full(d, 0.15), ref, "all", False, check_pd=False)
    N = 30_011                                  # not a multiple of the point tile
    S = np.random.default_rng(3).uniform(-1, 1, (N, 2 * d))
    A, B = S[:, :d].copy(), S[:, d:].copy()
    sa = SobolIndices(d, info, nMC=N, A=A, B=B)
    got = sa.analyze(first=True, total=True, second=True)
    s1, st, s2 = reference_indices(A, B, ref, True)
    assert np.max(np.abs(got["first"] - s1)) < 1e-9 and np.max(np.abs(got["total"] - st)) < 1e-9
    assert set(got["second"]) == set(s2)
    assert max(abs(got["second"][k] - s2[k]) for k in s2) < 1e-9
    # x1 and x2 carry the variance, x4 and x5 none
    assert got["first"][0] > 0.3 and got["first"][1] > 0.03 and abs(got["first"][3]) < 0.02 and got["total"][4] < 0.02
    # the sums add over row shards (what every rank contributes under data parallelism)
    sets = [(i,) for i in range(d)]
    full = sobol_sums(A, B, info, sets)
    cut = 12_345
    p1, p2 = sobol_sums(A[:cut], B[:cut], info, sets), sobol_sums(A[cut:], B[cut:], info, sets)
    for f, a, b in zip(full, p1, p2):
        assert np.allclose(f, a + b, rtol=1e-12, atol=0)
    # only total order: first stays zero, as calc_ft_order(False, True) returns it
    only_t = SobolIndices(d, info, nMC=N, A=A, B=B).analyze(first=False, total=True)
    assert np.all(only_t["first"] == 0) and np.max(np.abs(only_t["total"] - st)) < 1e-9
    with pytest.raises(NotImplementedError):
        SobolIndices(d, None, problem="hidimenra")
