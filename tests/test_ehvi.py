"""2-objective EHVI.

CPU: oracle/ehvi_oracle.py against tests/golden/ehvi2d.npz (outputs of the live reference's
kadal/optim_tools/ehvi/exi2d.py) — bit-identical — and, when /root/reference is present, against the
live function itself.  GPU: the CUDA path (kb200_ehvi2d through the C ABI) against the golden
vectors and the oracle.  exi2d is not covered by the north star's tolerances; the CUDA cells differ
from the reference only through the last ulp of erf/exp, the bound used here is 1e-12 relative to
max(|ref|, 1e-3 * box area)."""
import copy

import numpy as np
import pytest

from conftest import load_golden
from oracle import ehvi_oracle as eo, kriging_oracle as ko, ref_loader

TOL_EHVI = 1e-12


def golden_cases():
    g = load_golden("ehvi2d.npz")
    return [(g["P%d" % c], g["r%d" % c], g["mu%d" % c], g["s%d" % c], g["v%d" % c]) for c in range(int(g["ncases"]))]


def test_oracle_matches_reference_golden_bit_for_bit():
    for P, r, mu, s, v in golden_cases():
        got = eo.exi2d_batch(P, r, mu, s)
        assert np.array_equal(got, v)
        assert np.all(v >= 0) and v.max() > 0


@pytest.mark.skipif(not ref_loader.available(), reason="live reference not present")
def test_oracle_matches_live_reference():
    import warnings
    ref_loader._ensure_path()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from kadal.optim_tools.ehvi.exi2d import exi2d
    rng = np.random.default_rng(5)
    for k in (1, 3, 8):
        a, b = np.sort(rng.uniform(0, 2, k)), np.sort(rng.uniform(0, 2, k))[::-1]
        P = np.c_[a, b][rng.permutation(k)]
        r = np.array([2.5, 2.2])
        for _ in range(4):
            mu, s = rng.uniform(-0.5, 2.6, 2), rng.uniform(1e-4, 1.0, 2)
            assert eo.exi2d(P, r, mu, s) == exi2d(P, r, mu, s)
    # dominated points in the "front" and a degenerate spread behave like the reference too
    P = np.array([[0.2, 0.8], [0.5, 0.9], [0.7, 0.1]])
    for s in ([0.1, 0.2], [0.0, 0.2]):
        with np.errstate(all="ignore"):
            a, b = eo.exi2d(P, [1.0, 1.0], [0.4, 0.4], np.array(s)), exi2d(P, [1.0, 1.0], np.array([0.4, 0.4]), np.array(s))
        assert a == b or (np.isnan(a) and np.isnan(b))


def _err(got, ref, r):
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * r[0] * r[1])))


@pytest.mark.gpu
def test_cuda_exi2d_matches_golden_and_oracle():
    from kadal_b200 import ehvi
    for P, r, mu, s, v in golden_cases():
        got = ehvi.exi2d_batch(P, r, mu, s)
        assert _err(got, v, r) < TOL_EHVI
        assert abs(ehvi.exi2d(P, r, mu[0], s[0]) - v[0]) <= TOL_EHVI * max(abs(v[0]), 1e-3)
        assert np.array_equal(ehvi.exi2d_batch(P, r, mu, s, sign=-1.0), -got)
    # a large front (k = 300: numpy's pairwise recursion beyond 128 column sums) and many candidates
    rng = np.random.default_rng(9)
    k = 300
    P = np.c_[np.sort(rng.uniform(0, 1, k)), np.sort(rng.uniform(0, 1, k))[::-1]][rng.permutation(k)]
    r = np.array([1.1, 1.05])
    mu, s = rng.uniform(0, 1.1, (3000, 2)), rng.uniform(1e-3, 0.3, (3000, 2))
    got = ehvi.exi2d_batch(P, r, mu, s)
    idx = rng.choice(3000, 12, replace=False)
    T = eo.cell_table(P, r)
    ref = np.array([eo.exi2d(P, r, mu[i], s[i], T) for i in idx])
    assert _err(got[idx], ref, r) < TOL_EHVI
    assert np.all(got >= 0) and np.all(np.isfinite(got))
    # empty front: the whole box below r is improvement
    e0 = ehvi.exi2d_batch(np.zeros((0, 2)), r, mu[:5], s[:5])
    ref0 = eo.exi2d_batch(np.zeros((0, 2)), r, mu[:5], s[:5])
    assert _err(e0, ref0, r) < TOL_EHVI
    # the reference's degenerate inputs propagate the same way (zero spread -> NaN)
    with np.errstate(all="ignore"):
        bad = ehvi.exi2d_batch(P[:3], r, mu[:1], np.array([[0.0, 0.1]]))
        badr = eo.exi2d(P[:3], r, mu[0], np.array([0.0, 0.1]))
    assert (np.isnan(bad[0]) and np.isnan(badr)) or abs(bad[0] - badr) < 1e-12
    with pytest.raises(NotImplementedError):
        ehvi.ehvicalc_vec(mu[:2], P, {"refpoint": r}, [None, None, None])


@pytest.mark.gpu
def test_ehvicalc_vec_scores_a_population_like_the_reference():
    """EHVIcomputation.py:169-175: pred/SSqr of two objective models for a whole population, then
    -exi2d per candidate (the variance goes where a std-dev is meant, as in the reference)."""
    import kadal_b200
    from kadal_b200 import ehvi
    rng = np.random.default_rng(31)
    n, d = 90, 4
    X = rng.uniform(-1, 1, (n, d))
    y1 = ((X[:, 0] - 0.3) ** 2 + X[:, 1] ** 2)[:, None] + 1.0
    y2 = ((X[:, 0] + 0.4) ** 2 + (X[:, 2] - 0.2) ** 2)[:, None] + 1.0
    infos, refs = [], []
    for y in (y1, y2):
        info = ko.make_kriginfo(X, y)
        kadal_b200.likelihood(np.full(d, 0.1), info, "all", False)
        r = copy.deepcopy(info)
        ko.likelihood(np.full(d, 0.1), r, "all", False, check_pd=False)
        infos.append(info)
        refs.append(r)
    Y = np.c_[y1, y2]
    nd = [i for i in range(n) if not np.any(np.all(Y <= Y[i], axis=1) & np.any(Y < Y[i], axis=1))]
    y_par = Y[nd]
    mobo = {"refpoint": np.array([Y[:, 0].max() * 1.1, Y[:, 1].max() * 1.1])}
    xpop = rng.uniform(-1, 1, (64, d))
    hv = ehvi.ehvicalc_vec(xpop, y_par, mobo, infos)
    pr = [ko.prediction(xpop, r, ["pred", "SSqr"]) for r in refs]
    pred = np.c_[pr[0][0], pr[1][0]]
    ssqr = np.c_[pr[0][1], pr[1][1]]
    ref = eo.ehvicalc_vec(pred, ssqr, y_par, mobo["refpoint"])
    assert hv.shape == (64,) and np.all(hv <= 0)
    # the cell integrals amplify the predictor's 1e-8 / 1e-7 tolerances through (a - mu)/s
    assert np.max(np.abs(hv - ref)) < 1e-6 * np.max(np.abs(ref))
    assert int(np.argmin(hv)) == int(np.argmin(ref))
    one = ehvi.ehvicalc(xpop[3], y_par, mobo, infos)
    assert one == hv[3] or (hv[3] == 0 and one > 0)
