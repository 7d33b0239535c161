"""3-objective EHVI, the reference's KMAC C++ extension (second half).

CPU: oracle/ehvi3d_oracle.py against the reference's own known answers (test_single.py:21, test_multi.py:26-27),
against tests/golden/ehvi3d.npz (outputs of the reference's C++ compiled by oracle/kmac_ref/Makefile) and, when
oracle/_ref/libkmac_ref.so is present, against that library live.  GPU: kb200_ehvi3d through the C ABI against the
same.  KMAC is not covered by the north star's tolerances; the CUDA terms differ from the reference's only through
the last ulp of erf / exp, the bound used is 1e-11 relative to max(|ref|, 1e-3 x the volume of the box between the
reference point and the front's upper corner) -- a cell is a difference of terms of the size of that box, so this
is a few hundred ulps of them, the same convention as the 2-objective tests."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import ehvi3d_oracle as eo, kmac_ref_loader as kr

TOL = 1e-11

_spec = importlib.util.spec_from_file_location("make_golden_ehvi3d", os.path.join(GOLDEN, "make_golden_ehvi3d.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


def golden_cases():
    g = load_golden("ehvi3d.npz")
    return [tuple(g["%s%d" % (k, c)] for k in ("P", "r", "mu", "s", "v", "w")) for c in range(int(g["ncases"]))]


def err(got, ref, P, r):
    box = np.prod((P.max(axis=0) if len(P) else r + 10.0) - r)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * box)))


def test_oracle_reproduces_the_references_known_answers():
    v = eo.ehvi3d_multi(mk.KAT_SINGLE_FRONT, mk.KAT_REF, mk.KAT_SINGLE_MEAN, np.zeros((1, 3)))
    assert np.isclose(v[0], mk.KAT_SINGLE_ANSWER)                                  # test_single.py:21
    v = eo.ehvi3d_multi(mk.KAT_MULTI_FRONT, mk.KAT_REF, mk.KAT_MULTI_MEAN, np.zeros((3, 3)))
    assert np.isclose(v, mk.KAT_MULTI_ANSWER).all()                                # test_multi.py:26-27
    # the third answer of test_multi.py is an artefact of the early exit: alone, that candidate improves the front
    w = eo.ehvi3d_multi(mk.KAT_MULTI_FRONT, mk.KAT_REF, mk.KAT_MULTI_MEAN, np.zeros((3, 3)), independent=True)
    assert np.isclose(w[:2], mk.KAT_MULTI_ANSWER[:2]).all() and w[2] > 3.9


def test_oracle_matches_golden():
    cs = golden_cases()
    assert np.isclose(cs[0][4][0], mk.KAT_SINGLE_ANSWER) and np.isclose(cs[1][4], mk.KAT_MULTI_ANSWER).all()
    for c, (P, r, mu, s, v, w) in enumerate(cs):
        if len(P) > 20:
            mu, s, v, w = mu[:24], s[:24], v[:24], w[:24]       # a prefix of a batch is scored like the batch
        assert err(eo.ehvi3d_multi(P, r, mu, s), v, P, r) < TOL, c
        assert err(eo.ehvi3d_multi(P, r, mu, s, independent=True), w, P, r) < TOL, c
    # inputs of the generator are reproducible from the seed
    for (P, r, mu, s), (P2, r2, mu2, s2, _, _) in zip(mk.cases(), cs):
        assert np.array_equal(P, P2) and np.array_equal(mu, mu2) and np.array_equal(s, s2)


def test_front_filter_is_the_wrappers():
    P = np.array([[1., 1, 1], [2, 2, 2], [0, 0, 0], [3, 0, 0], [2, 2, 2]])
    F = eo.kmac_front(P)
    # a new point removes earlier ones it dominates or equals; it is never rejected itself (kmac_wrapper.cpp:24-34)
    assert np.array_equal(F, np.array([[3., 0, 0], [2, 2, 2]]))
    assert np.array_equal(eo.kmac_front(np.array([[2., 2, 2], [1, 1, 1]])), np.array([[2., 2, 2], [1, 1, 1]]))


@pytest.mark.skipif(not kr.available(), reason="oracle/_ref/libkmac_ref.so not built")
def test_oracle_matches_the_compiled_reference_live():
    rng = np.random.default_rng(77)
    for k in (0, 1, 4, 9, 15):
        P = mk.random_front(rng, k) if k else np.zeros((0, 3))
        r = np.array([-10.0, -10.5, -11.0])
        mu, s = rng.uniform(-10.5, -0.5, (21, 3)), rng.uniform(0.05, 3.0, (21, 3))
        assert err(eo.ehvi3d_multi(P, r, mu, s), kr.ehvi3d_multi(P, r, mu, s), P, r) < TOL
        alone = np.array([kr.ehvi3d_multi(P, r, mu[i:i + 1], s[i:i + 1])[0] for i in range(len(mu))])
        assert err(eo.ehvi3d_multi(P, r, mu, s, independent=True), alone, P, r) < TOL


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_ehvi3d_matches_reference_golden_and_known_answers():
    from kadal_b200 import ehvi
    for c, (P, r, mu, s, v, w) in enumerate(golden_cases()):
        got = ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s)
        assert err(got, v, P, r) < TOL, c
        alone = ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s, mode="independent")
        assert err(alone, w, P, r) < TOL, c
        assert np.array_equal(ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s, sign=-1.0), -got)
        # independent mode does not depend on the batch composition; reference mode does (that is the reference)
        perm = np.random.default_rng(c).permutation(len(mu))
        assert np.array_equal(ehvi.ehvi3d_sliceupdate_multi(P, r, mu[perm], s[perm], mode="independent"), alone[perm])
    v = ehvi.ehvi3d_sliceupdate_multi(mk.KAT_SINGLE_FRONT, mk.KAT_REF, mk.KAT_SINGLE_MEAN, np.zeros((1, 3)))
    assert np.isclose(v[0], mk.KAT_SINGLE_ANSWER)
    v = ehvi.ehvi3d_sliceupdate_multi(mk.KAT_MULTI_FRONT, mk.KAT_REF, mk.KAT_MULTI_MEAN, np.zeros((3, 3)))
    assert np.isclose(v, mk.KAT_MULTI_ANSWER).all()
    with pytest.raises(ValueError):
        ehvi.ehvi3d_sliceupdate_multi(mk.KAT_MULTI_FRONT, mk.KAT_REF[:2], mk.KAT_MULTI_MEAN, np.zeros((3, 3)))


def tied_front(rng, k):
    """Integer-valued objectives: many equal coordinates (zero-length cells, ties in the three sort orders)."""
    P = rng.integers(-9, 0, (40, 3)).astype(float)
    nd = [i for i in range(len(P)) if not any((P[j] >= P[i]).all() and (P[j] > P[i]).any() for j in range(len(P)))]
    P = np.unique(P[nd], axis=0)[:k]
    return P[rng.permutation(len(P))]


@pytest.mark.skipif(not kr.available(), reason="oracle/_ref/libkmac_ref.so not built")
def test_oracle_matches_the_compiled_reference_on_tied_coordinates():
    rng = np.random.default_rng(11)
    r = np.array([-10.0, -10.0, -10.0])
    for _ in range(6):
        P = tied_front(rng, int(rng.integers(2, 12)))
        mu, s = rng.uniform(-10, -1, (15, 3)), rng.uniform(0.2, 2, (15, 3))
        assert err(eo.ehvi3d_multi(P, r, mu, s), kr.ehvi3d_multi(P, r, mu, s), P, r) < TOL


@pytest.mark.gpu
def test_cuda_ehvi3d_tied_coordinates():
    from kadal_b200 import ehvi
    rng = np.random.default_rng(12)
    r = np.array([-10.0, -10.0, -10.0])
    for _ in range(6):
        P = tied_front(rng, int(rng.integers(2, 12)))
        mu, s = rng.uniform(-10, -1, (33, 3)), rng.uniform(0.2, 2, (33, 3))
        mu[4] = P[0]                                   # a mean exactly on a (tied) grid line
        for mode, indep in (("reference", False), ("independent", True)):
            got = ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s, mode=mode)
            assert err(got, eo.ehvi3d_multi(P, r, mu, s, independent=indep), P, r) < TOL
            if kr.available() and not indep:
                assert err(got, kr.ehvi3d_multi(P, r, mu, s), P, r) < TOL


@pytest.mark.gpu
def test_cuda_ehvi3d_large_front_and_population():
    """BASELINE config 5's scoring shape: 1500 candidates (more than the reference's static tables hold: 1000)
    over a 120-point front; checked against the oracle on a prefix and, when the compiled reference is there,
    against it on the first 1000."""
    from kadal_b200 import ehvi
    rng = np.random.default_rng(5)
    P = mk.random_front(rng, 120)
    r = np.array([-10.0, -10.0, -10.0])
    mu, s = rng.uniform(-10.2, -0.5, (1500, 3)), rng.uniform(0.3, 2.5, (1500, 3))
    got = ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s)
    alone = ehvi.ehvi3d_sliceupdate_multi(P, r, mu, s, mode="independent")
    assert got.shape == (1500,) and np.all(np.isfinite(got)) and np.all(got >= 0) and np.all(alone >= got - 1e-9)
    assert err(ehvi.ehvi3d_sliceupdate_multi(P, r, mu[:6], s[:6]), got[:6], P, r) < TOL        # prefix property
    assert err(got[:6], eo.ehvi3d_multi(P, r, mu[:6], s[:6]), P, r) < TOL
    if kr.available():
        assert err(got[:1000], kr.ehvi3d_multi(P, r, mu[:1000], s[:1000]), P, r) < TOL
        idx = rng.choice(1500, 8, replace=False)
        ref = np.array([kr.ehvi3d_multi(P, r, mu[i:i + 1], s[i:i + 1])[0] for i in idx])
        assert err(alone[idx], ref, P, r) < TOL


@pytest.mark.gpu
def test_ehvicalc_kmac3d_scores_a_population_like_the_reference():
    """EHVIcomputation.py:189-251: pred / SSqr of three objective models for a population, negated into KMAC's
    maximisation convention, the variance where a standard deviation is meant (:235)."""
    import copy
    import kadal_b200
    from kadal_b200 import ehvi
    from oracle import kriging_oracle as ko
    rng = np.random.default_rng(33)
    n, d = 80, 3
    X = rng.uniform(-1, 1, (n, d))
    ys = [((X[:, 0] - 0.3) ** 2 + X[:, 1] ** 2)[:, None] + 1.0, ((X[:, 0] + 0.4) ** 2 + (X[:, 2] - 0.2) ** 2)[:, None] + 1.0,
          ((X[:, 1] - 0.5) ** 2 + (X[:, 2] + 0.3) ** 2)[:, None] + 1.0]
    infos, refs = [], []
    for y in ys:
        info = ko.make_kriginfo(X, y)
        kadal_b200.likelihood(np.full(d, 0.1), info, "all", False)
        rr = copy.deepcopy(info)
        ko.likelihood(np.full(d, 0.1), rr, "all", False, check_pd=False)
        infos.append(info), refs.append(rr)
    Y = np.hstack(ys)
    nd = [i for i in range(n) if not np.any(np.all(Y <= Y[i], axis=1) & np.any(Y < Y[i], axis=1))]
    y_par = Y[nd]
    mobo = {"refpoint": Y.max(axis=0) * 1.1}
    xpop = rng.uniform(-1, 1, (48, d))
    pr = [ko.prediction(xpop, rr, ["pred", "SSqr"]) for rr in refs]
    pred, ssqr = np.hstack([p[0] for p in pr]), np.hstack([p[1] for p in pr])
    for mode, indep in (("reference", False), ("independent", True)):
        hv = ehvi.ehvicalc_kmac3d(xpop, y_par, mobo, infos, mode=mode)
        ref = -eo.ehvi3d_multi(-y_par, -mobo["refpoint"], -pred, ssqr, independent=indep)
        assert hv.shape == (48,) and np.all(hv <= 0)
        # the cell integrals amplify the predictor's 1e-8 / 1e-7 tolerances through (c - mu) / s
        assert np.max(np.abs(hv - ref)) < 1e-6 * max(np.max(np.abs(ref)), 1e-12)
    one = ehvi.ehvicalc_kmac3d(xpop[0], y_par, mobo, infos)
    assert np.isscalar(one) or np.ndim(one) == 0
