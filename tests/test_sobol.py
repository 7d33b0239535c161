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


def _fit_c4():
    import kadal_b200
    w = ko.workload("C4")
    info = ko.make_kriginfo(w["X"], w["y"])
    kadal_b200.likelihood(w["theta"].copy(), info, "all", False)
    return info


def test_device_sobol_generator_is_bit_identical_to_the_host_one():
    """kb200_sobol_points_dev against kadal_b200.sampling.sampling('sobolnew', ...) (scipy's unscrambled generator when
    the sobolsampling package is absent): same direction numbers, closed-form Gray-code xor instead of the sequential
    recurrence, realval scaling without FMA -- every bit equal, at an offset too (a rank's row shard)."""
    import torch
    from kadal_b200 import sampling as sp
    if sp.sobol_direction_numbers(4, 10) is None:
        pytest.skip("the sobolsampling package is installed: its tables are not exposed, A / B are drawn on the host")
    for dims, n in ((40, 70_001), (3, 1000), (1, 5)):
        lo, hi = np.linspace(-2.0, -1.0, dims), np.linspace(0.5, 3.0, dims)
        unit, real = sp.sampling("sobolnew", dims, n, result="real", upbound=hi, lobound=lo)
        out = torch.empty((n, dims), dtype=torch.float64, device="cuda:0")
        assert sp.sobol_points_device(n, dims, lo, hi, device=0, out_ptr=out.data_ptr())
        assert np.array_equal(out.cpu().numpy(), real)
        assert sp.sobol_points_device(n, dims, device=0, out_ptr=out.data_ptr())
        assert np.array_equal(out.cpu().numpy(), unit)
        k = n // 3
        part = torch.empty((n - k, dims), dtype=torch.float64, device="cuda:0")
        assert sp.sobol_points_device(n - k, dims, lo, hi, first_row=1 + k, device=0, out_ptr=part.data_ptr())
        assert np.array_equal(part.cpu().numpy(), real[k:])


def test_sobol_sums_device_resident_and_generated_equal_the_host_path():
    """kb200_sobol_sums_dev (A, B already in HBM) and kb200_sobol_sums_gen (A, B drawn on the device) against
    kb200_sobol_sums on the host-generated plan: the same chunks through the same kernels -> identical sums."""
    import torch
    import kadal_b200
    from kadal_b200 import sampling as sp
    from kadal_b200.sensitivity import SobolIndices, sobol_sums, sobol_sums_generated
    rng = np.random.default_rng(4)
    n, d = 120, 4
    X = rng.uniform(-1, 1, (n, d))
    y = (np.sin(2 * X[:, 0]) + X[:, 1] ** 2 + 0.2 * X[:, 3])[:, None] + 3.0
    lbm, ubm = np.full(d, -1.0), np.full(d, 1.0)
    info = ko.make_kriginfo(X, y, lb=lbm, ub=ubm)
    kadal_b200.likelihood(np.full(d, 0.1), info, "all", False)
    N = 300_017                                   # two chunks of 2^18 rows
    lo2, hi2 = np.r_[lbm, lbm], np.r_[ubm, ubm]
    _, S = sp.sampling("sobolnew", 2 * d, N, result="real", upbound=hi2, lobound=lo2)
    A, B = np.ascontiguousarray(S[:, :d]), np.ascontiguousarray(S[:, d:])
    sets = [(i,) for i in range(d)] + [(0, 2)]
    host = sobol_sums(A, B, info, sets)
    tA, tB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    torch.cuda.synchronize()
    dev = sobol_sums((tA.data_ptr(), N), (tB.data_ptr(), N), info, sets)
    for h, g in zip(host, dev):
        assert np.array_equal(h, g)
    gen = sobol_sums_generated(N, info, sets, lo2, hi2)
    if gen is not None:
        for h, g in zip(host, gen):
            assert np.array_equal(h, g)
        # the drop-in class draws on the device when no sample is given and gives the reference's indices
        got = SobolIndices(d, info, ub=hi2, lb=lo2, nMC=N).analyze(first=True, total=True)
        want = SobolIndices(d, info, nMC=N, A=A, B=B).analyze(first=True, total=True)
        assert np.array_equal(got["first"], want["first"]) and np.array_equal(got["total"], want["total"])


def test_c4_full_size_saltelli_sums():
    """BASELINE config 4 at its full size: n = 800, d = 20, N = 909 091 base rows, 22 prediction sets (A, B, 20 x C_i) =
    2.0e7 predictions in one kb200_sobol_sums_gen call, nothing but 44 doubles back.  Checked against chunked bulk
    predictions of the same device path reduced on the host with the reference's np.sum formulas (sobol_ind.py:84-85,
    136-139), and the Saltelli identities of a plan."""
    from kadal_b200 import sampling as sp
    from kadal_b200.prediction import predict_arrays
    from kadal_b200.sensitivity import sobol_sums, sobol_sums_generated
    info = _fit_c4()
    d, N = 20, 909_091
    lo2, hi2 = -np.ones(2 * d), np.ones(2 * d)
    sets = [(i,) for i in range(d)]
    gen = sobol_sums_generated(N, info, sets, lo2, hi2)
    _, S = sp.sampling("sobolnew", 2 * d, N, result="real", upbound=hi2, lobound=lo2)
    A, B = np.ascontiguousarray(S[:, :d]), np.ascontiguousarray(S[:, d:])
    host = sobol_sums(A, B, info, sets)
    if gen is None:
        gen = host
    for h, g in zip(host, gen):
        assert np.array_equal(h, g)
    ya = predict_arrays(A, info, ("pred",))["pred"]
    yb = predict_arrays(B, info, ("pred",))["pred"]
    base, yayc, ybyc = gen
    assert np.allclose(base, [ya.sum(), (ya ** 2).sum(), yb.sum(), (yb ** 2).sum()], rtol=1e-11, atol=0)
    for i in (0, 7, 19):
        C = B.copy()
        C[:, i] = A[:, i]
        yc = predict_arrays(C, info, ("pred",))["pred"]
        assert abs(yayc[i] - np.sum(ya * yc)) <= 1e-11 * abs(yayc[i]) and abs(ybyc[i] - np.sum(yb * yc)) <= 1e-11 * abs(ybyc[i])
    fo2 = (base[0] / N) ** 2
    den = base[1] / N - fo2
    s1 = (yayc / N - fo2) / den
    st = 1 - (ybyc / N - fo2) / den
    # y = sin(x0 + x1 + x2) + noise: the first three inputs carry the variance, the rest next to nothing
    assert s1[:3].min() > 0.05 and np.abs(s1[3:]).max() < 0.02 and np.all(st[:3] > s1[:3] - 1e-3)
