"""GPU: the CUDA path (through the C ABI) against the oracle, the golden vectors of the live
reference, and size-independent properties at BASELINE.json's full sizes.

Tolerances are the north star's: ln-likelihood 1e-9 relative, mean 1e-8, variance 1e-7
(denominators floored as in SURVEY.md Appendix E: max(|ref|, 1) for NLL / mean,
max(|ref|, 1e-6*sigma^2) for SSqr), arg-min candidate index exact.  For candidates with
cond(Psi) > 1e8 the reference's own reordering noise exceeds 1e-9 (Appendix E: 6-8e-10 at
cond 2e8); there the bound is scaled with the condition number and says so.
"""
import copy

import numpy as np
import pytest

from conftest import relerr
import cases
from oracle import kriging_oracle as ko

pytestmark = pytest.mark.gpu

TOL_NLL, TOL_MEAN, TOL_VAR = 1e-9, 1e-8, 1e-7


@pytest.fixture(scope="module")
def kb():
    import kadal_b200
    from kadal_b200 import _capi
    _capi.require_gpu()   # loud failure if the CUDA extension / device is missing
    return kadal_b200


def nll_tol(cond):
    """1e-9, or half of cond * ulp where that is larger: SURVEY.md Appendix E measures the
    reference's own reordering noise at 6-8e-10 for cond 2e8 (about 0.03 cond*ulp) and a
    one-ulp change of the kernel entries at up to 0.07 cond*ulp."""
    return max(TOL_NLL, 0.5 * cond * 2.0 ** -53)


def conds_of(info, pop):
    """cond(Psi + eps I) per candidate, from the oracle (the checker)."""
    out = []
    for x in pop:
        A = copy.deepcopy(info)
        ko.likelihood(np.array(x, dtype=float), A, "all", False, check_pd=False)
        n = A["Psi"].shape[0]
        out.append(np.linalg.cond(A["Psi"] + np.eye(n) * 10.0 ** float(A["nugget"])))
    return np.array(out)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", cases.APPENDIX_C_CASES)
def test_appendix_c_against_reference_golden(kb, golden_c, name):
    g = golden_c
    info, x, tv = cases.appendix_c_info(g, name)
    v = kb.likelihood(x.copy(), copy.deepcopy(info), "default", tv)
    assert relerr(v, g[name + "_nll"]) < TOL_NLL
    out = kb.likelihood(x.copy(), info, "all", tv)
    assert out is info
    s2 = float(g[name + "_s2"][0])
    assert relerr(info["NegLnLike"], g[name + "_nll"]) < TOL_NLL
    assert relerr(info["BE"], g[name + "_BE"]) < TOL_MEAN
    assert relerr(np.asarray(info["SigmaSqr"]).ravel(), g[name + "_s2"], 1e-300) < TOL_MEAN
    assert relerr(info["U"], g[name + "_U"]) < TOL_MEAN
    assert relerr(info["Psi"], g[name + "_Psi"], 1e-300) < 1e-13
    assert info["BE"].shape == (1, 1)
    assert np.shape(info["SigmaSqr"]) == (() if tv else (1, 1))
    p = kb.prediction(g["xp"], info, cases.PT)
    assert relerr(p[0].ravel(), g[name + "_pred"]) < TOL_MEAN
    assert relerr(p[1].ravel(), g[name + "_SSqr"], 1e-6 * s2) < TOL_VAR
    assert relerr(p[2].ravel(), g[name + "_s"], 1e-3 * np.sqrt(s2)) < TOL_VAR
    assert relerr(p[4].ravel(), g[name + "_PoI"], 1e-12) < 1e-6
    ei_ref = g[name + "_EI"]
    big = np.abs(ei_ref) > 1e-200
    assert relerr(p[3].ravel()[big], ei_ref[big], 1e-300) < 1e-4
    for a in p:
        assert a.shape == (3, 1)


def test_c3_population_and_bulk_predict_against_golden(kb, golden_c3):
    g = golden_c3
    info, xp = cases.c3_info(g)
    nll, inf = kb.likelihood_batch(g["pop"], info, False)
    assert np.all(inf == 0)
    conds = conds_of(info, g["pop"])
    rel = np.abs(nll - g["pop_nll"]) / np.maximum(np.abs(g["pop_nll"]), 1.0)
    tol = np.array([nll_tol(c) for c in conds])
    assert np.all(rel < tol), (rel / tol).max()
    assert np.all(rel[conds < 1e7] < TOL_NLL)                       # plain 1e-9 wherever cond*ulp allows it
    assert int(np.argmin(nll)) == int(np.argmin(g["pop_nll"]))
    kb.likelihood(g["fit_x"].copy(), info, "all", False)
    s2 = float(g["fit_s2"][0])
    p = kb.prediction(xp, info, cases.PT + ["PoF"])
    assert relerr(p[0].ravel(), g["pred"]) < TOL_MEAN
    assert relerr(p[1].ravel(), g["SSqr"], 1e-6 * s2) < TOL_VAR
    assert relerr(p[2].ravel(), g["s"], 1e-3 * np.sqrt(s2)) < TOL_VAR
    assert relerr(p[4].ravel(), g["PoI"], 1e-9) < 1e-5
    assert relerr(p[5].ravel(), g["PoF"], 1e-9) < 1e-5
    # same model, predictor imported from the *reference's* factor (KrigInfo['U'] route)
    ref = copy.deepcopy(cases.c3_info(g)[0])
    ko.likelihood(g["fit_x"].copy(), ref, "all", False, check_pd=False)
    q = kb.prediction(xp, ref, ["pred", "SSqr"])
    assert relerr(q[0].ravel(), g["pred"]) < TOL_MEAN
    assert relerr(q[1].ravel(), g["SSqr"], 1e-6 * s2) < TOL_VAR


def test_std_normalisation_against_golden(kb, golden_std):
    g = golden_std
    info = cases.std_info(g)
    kb.likelihood(g["x"].copy(), info, "all", False)
    s2 = float(g["s2"][0])
    assert relerr(info["NegLnLike"], g["nll"]) < TOL_NLL
    p = kb.prediction(g["xp"], info, ["pred", "SSqr", "s"])
    assert relerr(p[0].ravel(), g["pred"]) < TOL_MEAN
    assert relerr(p[1].ravel(), g["SSqr"], 1e-6 * s2) < TOL_VAR


# ------------------------------------------------------------------ round-2 option goldens (r2_options.npz)
@pytest.mark.parametrize("order", [1, 2])
def test_trend_order_likelihood_against_reference_golden(kb, golden_r2, order):
    """TrendOrder 1 (p = 4 trend columns) and 2 (p = 6): F from the reference's compute_regression_mat
    (trendfunction.py:72-111) through the fused forward substitution of [y F] and the p x p generalised least squares
    of likelihood_func.py:183-193; concentrated and trainvar likelihood, population + 'all' call.  Prediction with
    p > 1 raises in the reference (prediction.py:250) and here."""
    g, t = golden_r2, "trend%d_" % order
    info = cases.trend_info(g, order)
    p = g[t + "F"].shape[1]
    nll, inf = kb.likelihood_batch(g[t + "pop"], info, False)
    assert np.all(inf == 0)
    conds = conds_of(info, g[t + "pop"])
    rel = np.abs(nll - g[t + "pop_nll"]) / np.maximum(np.abs(g[t + "pop_nll"]), 1.0)
    # per candidate: inside the bound, or closer to the binary128 value of the formula than the reference itself is
    # (with p trend columns the p x p normal equations add their own conditioning: the reference is up to 7e-9 off)
    den = np.maximum(np.abs(g[t + "pop_truth"]), 1.0)
    err, err_ref = np.abs(nll - g[t + "pop_truth"]) / den, np.abs(g[t + "pop_nll"] - g[t + "pop_truth"]) / den
    ok = (rel < np.array([nll_tol(c) for c in conds])) | (err <= err_ref)
    assert np.all(ok), (rel[~ok], err[~ok], err_ref[~ok])
    assert err.max() <= max(err_ref.max(), TOL_NLL)
    assert np.all((rel[conds < 1e7] < TOL_NLL) | (err <= err_ref)[conds < 1e7])
    assert int(np.argmin(nll)) == int(np.argmin(g[t + "pop_nll"]))
    A = copy.deepcopy(info)
    kb.likelihood(g[t + "x"].copy(), A, "all", False)
    tr = float(g[t + "nll_truth"])
    assert (relerr(A["NegLnLike"], g[t + "nll"]) < TOL_NLL
            or abs(A["NegLnLike"] - tr) <= abs(float(g[t + "nll"]) - tr)), (A["NegLnLike"], float(g[t + "nll"]), tr)
    assert A["BE"].shape == (p, 1) and relerr(A["BE"], g[t + "BE"], 1e-6 * np.abs(g[t + "BE"]).max()) < 1e-7
    assert relerr(np.asarray(A["SigmaSqr"]).ravel(), g[t + "s2"], 1e-300) < TOL_MEAN
    assert relerr(A["U"], g[t + "U"]) < TOL_MEAN
    B = copy.deepcopy(info)
    kb.likelihood(g[t + "tv_x"].copy(), B, "all", True)
    assert relerr(B["NegLnLike"], g[t + "tv_nll"]) < TOL_NLL
    assert relerr(B["BE"], g[t + "tv_BE"], 1e-6 * np.abs(g[t + "tv_BE"]).max()) < 1e-7
    assert np.shape(B["SigmaSqr"]) == () and relerr(B["SigmaSqr"], g[t + "tv_s2"][0]) < 1e-15
    with pytest.raises(NotImplementedError):
        kb.prediction(g[t + "Xraw"][:3], A, "pred")


def test_more_than_seven_trend_columns_fail_loudly(kb, golden_r2):
    """The fused forward substitution carries [y F] as one 8-row DMMA strip: p <= 7 (TrendOrder 2 at d = 3 is p = 10).
    No silent CPU diversion: the model constructor refuses."""
    from kadal_b200 import _capi
    g = golden_r2
    info = cases.trend_info(g, 1)
    info["F"] = np.hstack([info["F"], info["F"], info["F"][:, :2]])     # p = 10
    with pytest.raises(_capi.Kb200Error, match="trend columns"):
        kb.likelihood_batch(g["trend1_pop"][:2], info, False)


def test_iso_gaussian_against_reference_golden(kb, golden_r2):
    """kernel ['iso_gaussian'] (likelihood_func.py:60-61, KB200_FLAG_ISO): one length scale for every dimension,
    scalar x as minimize_scalar passes it; Theta written back as nvar copies; prediction from that state."""
    g = golden_r2
    info = cases.iso_info(g)
    out = kb.likelihood(float(g["iso_x"]), info, "all", False)
    assert out is info and relerr(info["NegLnLike"], g["iso_nll"]) < TOL_NLL
    assert np.array_equal(np.asarray(info["Theta"], dtype=float), g["iso_Theta"])
    s2 = float(g["iso_s2"][0])
    assert relerr(info["BE"], g["iso_BE"]) < TOL_MEAN and relerr(np.asarray(info["SigmaSqr"]).ravel(), g["iso_s2"]) < TOL_MEAN
    pr = kb.prediction(g["iso_xp"], info, ["pred", "SSqr", "s"])
    assert relerr(pr[0].ravel(), g["iso_pred"]) < TOL_MEAN
    assert relerr(pr[1].ravel(), g["iso_SSqr"], 1e-6 * s2) < TOL_VAR
    assert relerr(pr[2].ravel(), g["iso_s"], 1e-3 * np.sqrt(s2)) < TOL_VAR
    nll, inf = kb.likelihood_batch(g["iso_grid"][:, None], cases.iso_info(g), False)      # P x 1 population
    assert np.all(inf == 0) and relerr(nll, g["iso_grid_nll"]) < TOL_NLL
    assert relerr(kb.likelihood(float(g["iso_grid"][3]), cases.iso_info(g), "default", False), g["iso_grid_nll"][3]) < TOL_NLL


def test_pof_less_than_lcb_and_ebe_against_reference_golden(kb, golden_r2):
    """PoF with limittype '<' / '<=' (prediction.py:303-306), 'lcb' with its (npred, 1) - (1, npred) broadcast
    (prediction.py:268: an (npred, npred) array, variance not s) and the untransposed 'ebe' (:270)."""
    g = golden_r2
    info = cases.pof_lcb_info(g)
    kb.likelihood(g["pl_x"].copy(), info, "all", False)
    for lt, tag in (("<", "lt"), ("<=", "le"), (">", "gt")):
        info["limit"], info["limittype"] = float(g["pl_limit"]), lt
        pof = kb.prediction(g["pl_xp"], info, "PoF")
        assert pof.shape == (25, 1) and relerr(pof.ravel(), g["pl_pof_" + tag], 1e-9) < 1e-6
    info["limittype"] = "=="
    with pytest.raises(ValueError):
        kb.prediction(g["pl_xp"], info, "PoF")
    info["sigmalcb"] = float(g["pl_sigmalcb"])
    lcb = kb.prediction(g["pl_xp"], info, "LCB")
    assert lcb.shape == (25, 25) and relerr(lcb, g["pl_lcb"]) < TOL_VAR
    ebe = kb.prediction(g["pl_xp"], info, "ebe")
    assert ebe.shape == (1, 25) and relerr(ebe, g["pl_ebe"], 1e-6 * float(np.asarray(info["SigmaSqr"]).ravel()[0])) < TOL_VAR
    del info["sigmalcb"]
    with pytest.raises(KeyError):           # nothing sets KrigInfo['sigmalcb'] in the reference either
        kb.prediction(g["pl_xp"], info, "LCB")


@pytest.mark.parametrize("kernel", [["gaussian"], ["iso_gaussian"]])
def test_single_hyperparameter_training_path(kb, kernel):
    """nbhyp == 1 (1-D input, or iso_gaussian at any d): the reference calls minimize_scalar(method='golden',
    bounds=...) (kriging_model.py:210-213), which scipy >= 1.5 rejects -- here five batched grids.
    The optimum must be at least as good as a dense oracle grid allows, and the fitted model must predict."""
    rng = np.random.default_rng(8)
    d = 1 if kernel == ["gaussian"] else 3
    n = 40
    X = rng.uniform(-3, 3, (n, d))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * X[:, :1] ** 2
    K = dict(X=X, y=y, nvar=d, nsamp=n, nrestart=3, ub=np.full(d, 3.0), lb=np.full(d, -3.0), kernel=list(kernel),
             nugget=-6, optimizer="lbfgsb")
    k = kb.Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    k.train(disp=False)
    assert k.nbhyp == 1 and k.train_stats["batches"] == 5
    best = float(np.asarray(k.KrigInfo["Theta"]).ravel()[0])
    assert np.asarray(k.KrigInfo["Theta"]).shape == (d,)
    grid = np.linspace(-5, 5, 65)            # the first of the five batched grids: the result can only be better
    probe = copy.deepcopy(k.KrigInfo)
    ref = np.array([ko.likelihood(float(t) if kernel == ["iso_gaussian"] else np.array([t]), probe, "default", False,
                                  check_pd=False) for t in grid])
    assert k.KrigInfo["NegLnLike"] <= ref.min() + 1e-9 * max(1.0, abs(ref.min()))
    assert abs(best - grid[int(np.argmin(ref))]) <= 10.0 / 64 + 1e-12
    R = copy.deepcopy(k.KrigInfo)
    ko.likelihood(float(best) if kernel == ["iso_gaussian"] else np.array([best]), R, "all", False, check_pd=False)
    xp = rng.uniform(-3, 3, (50, d))
    assert relerr(k.predict(xp, ["pred"]), ko.prediction(xp, R, ["pred"])) < TOL_MEAN


# ------------------------------------------------------------------ BASELINE config 2
@pytest.fixture(scope="module")
def c2():
    w = ko.workload("C2")
    return w, ko.make_kriginfo(w["X"], w["y"])


def test_c2_known_answer_and_anchors(kb, c2, golden_c2):
    w, info = c2
    v = kb.likelihood(w["kat_x"], copy.deepcopy(info), "default", False)
    assert abs(v - (-345.9939996418981)) < TOL_NLL * 346
    nA, infA = kb.likelihood_batch(w["popA"], info, False)
    nB, infB = kb.likelihood_batch(w["popB"], info, False)
    assert np.all(infA == 0) and np.all(infB == 0)
    assert relerr(nA[golden_c2["idxA"]], golden_c2["nllA"]) < TOL_NLL
    # popB reaches cond ~ 7e8: reference reordering noise ~ cond * ulp
    assert relerr(nB[golden_c2["idxB"]], golden_c2["nllB"]) < nll_tol(1e9)


def test_c2_full_population_properties(kb, c2):
    w, info = c2
    pop = w["popB"]
    nll, inf = kb.likelihood_batch(pop, info, False)
    assert nll.shape == (512,) and np.all(np.isfinite(nll))
    # independence of the batch composition: any sub-batch / permutation is bit-identical
    perm = np.random.default_rng(3).permutation(512)
    nll_p, _ = kb.likelihood_batch(pop[perm], info, False)
    assert np.array_equal(nll_p, nll[perm])
    # a single candidate takes the right-looking schedule (little parallel work), the population the
    # left-looking one: same value to rounding (section 4 of DESIGN.md), not the same bits
    one, _ = kb.likelihood_batch(pop[17:18], info, False)
    assert abs(one[0] - nll[17]) < 1e-9 * max(1.0, abs(nll[17]))
    # arg-min must match the oracle on the contenders (best 12 by GPU value + 12 random)
    rng = np.random.default_rng(5)
    idx = np.unique(np.r_[np.argsort(nll)[:12], rng.choice(512, 12, replace=False)])
    ref = ko.likelihood_population(pop[idx], copy.deepcopy(info), False, check_pd=False)
    assert idx[int(np.argmin(ref))] == int(np.argmin(nll))
    assert relerr(nll[idx], ref) < nll_tol(1e9)
    order_ref, order_gpu = np.argsort(ref), np.argsort(nll[idx])
    assert np.array_equal(order_ref[:6], order_gpu[:6])


def test_c2_every_candidate_against_live_reference(kb, c2, golden_c2_full, golden_c2_truth_full):
    """All 2 x 512 candidates of BASELINE config 2, through both factorisation schedules, against
      (a) the live reference's float64 result (tests/golden/c2_full.npz, make_golden.py c2full) and
      (b) the binary128 value of the reference's own formula on the same float64 Psi
          (tests/golden/c2_truth_full.npz, make_truth_full.py + oracle/truth/nll_quad.c).
    Per candidate:  rel(gpu, ref) <= 1e-9   OR   |gpu - truth| <= |ref - truth|,
    i.e. wherever the north star's 1e-9 is missed it is the *reference* that is further from the exact value of the
    formula (its float64 LAPACK path is 0.02 .. 1.4 cond ulp away at cond ~ 1e9).  Arg-min and the order of the best
    8 are exact.  The number of candidates in each branch goes to gpurun_out/parity_c2_counts.json (copied to
    profiles/ by tools/gpu_round2.sh)."""
    import json
    import os
    w, info = c2
    g, t = golden_c2_full, golden_c2_truth_full
    ulp = 2.0 ** -53
    counts = {}
    for pop in ("popA", "popB"):
        ref, cond, truth = g["nll_" + pop], g["cond_" + pop], t["truth_" + pop]
        assert np.all(np.isfinite(truth))
        den = np.maximum(1.0, np.abs(truth))
        err_ref = np.abs(ref - truth) / den
        nll, inf = kb.likelihood_batch(w[pop], info, False)                  # population: left-looking schedule
        assert np.all(inf == 0)
        assert int(np.argmin(nll)) == int(np.argmin(ref))
        assert np.array_equal(np.argsort(nll)[:8], np.argsort(ref)[:8])
        runs = {"left_looking_population": nll}
        if pop == "popB":
            # the 32 worst-conditioned candidates once more as a small batch: right-looking schedule
            sel = np.argsort(cond)[-32:]
            small, _ = kb.likelihood_batch(w[pop][sel], info, False)
            full = nll.copy()
            full[sel] = small
            runs["right_looking_worst32"] = full
        for name, got in runs.items():
            rel = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
            err = np.abs(got - truth) / den
            within = rel <= TOL_NLL
            closer = err <= err_ref
            bad = ~(within | closer)
            counts[pop + "/" + name] = {
                "candidates": int(len(ref)), "within_1e-9_of_reference": int(within.sum()),
                "beyond_1e-9_but_closer_to_binary128_truth_than_reference": int((~within & closer).sum()),
                "neither": int(bad.sum()), "max_rel_vs_reference": float(rel.max()),
                "max_err_vs_truth_gpu": float(err.max()), "max_err_vs_truth_reference": float(err_ref.max()),
                "max_err_vs_truth_gpu_in_cond_ulp": float((err / (cond * ulp)).max()),
                "max_err_vs_truth_reference_in_cond_ulp": float((err_ref / (cond * ulp)).max()),
                "median_err_vs_truth_gpu_in_cond_ulp": float(np.median(err / (cond * ulp))),
                "median_err_vs_truth_reference_in_cond_ulp": float(np.median(err_ref / (cond * ulp)))}
            assert not bad.any(), (pop, name, np.flatnonzero(bad)[:8], rel[bad][:8], err[bad][:8], err_ref[bad][:8])
            # and as a population the CUDA path is the more accurate of the two
            assert np.median(err) <= max(np.median(err_ref), 1e-15) and err.max() <= max(err_ref.max(), TOL_NLL)
    print("C2 parity counts:", json.dumps(counts))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(counts, open(os.path.join(out, "parity_c2_counts.json"), "w"), indent=1)


def test_c2_factor_round_trip(kb, c2):
    w, info = c2
    info = copy.deepcopy(info)
    x = w["popB"][3]
    kb.likelihood(x.copy(), info, "all", False)
    U, Psi = info["U"], info["Psi"]
    assert np.allclose(np.tril(U, -1), 0.0)
    A = Psi + np.eye(1000) * 1e-6
    assert np.linalg.norm(U.T @ U - A) / np.linalg.norm(A) < 1e-14
    assert np.array_equal(Psi, Psi.T) and np.all(np.diag(Psi) == 1.0)
    # Kriging interpolates the samples up to the nugget; variance there is ~ nugget-sized
    p, ss = kb.prediction(info["X"][:300], info, ["pred", "SSqr"])
    s2 = float(info["SigmaSqr"][0, 0])
    assert np.max(np.abs(p - info["y"][:300])) < 1e-3 * np.ptp(info["y"])
    assert np.max(np.abs(ss)) < 1e-4 * s2


# ------------------------------------------------------------------ shapes and edges
@pytest.mark.parametrize("n", [1, 2, 5, 127, 128, 129, 257, 300])
def test_tile_boundaries_likelihood_and_predict(kb, n):
    rng = np.random.default_rng(n)
    d = 3
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X.sum(1, keepdims=True)) + 3.0
    info = ko.make_kriginfo(X, y)
    x = np.array([0.0, 0.2, -0.1])
    A, B = copy.deepcopy(info), copy.deepcopy(info)
    ko.likelihood(x.copy(), A, "all", False, check_pd=False)
    kb.likelihood(x.copy(), B, "all", False)
    if n > 1:   # n = 1: sigma^2 = 0 exactly in exact arithmetic, ln(0) on both sides
        assert relerr(B["NegLnLike"], A["NegLnLike"]) < TOL_NLL
    assert relerr(B["U"], A["U"]) < TOL_MEAN
    for npred in (1, 127, 129):
        xp = rng.uniform(-1, 1, (npred, d))
        pa = ko.prediction(xp, A, ["pred", "SSqr"])
        pb = kb.prediction(xp, B, ["pred", "SSqr"])
        s2 = float(np.asarray(A["SigmaSqr"]).ravel()[0])
        assert relerr(pb[0], pa[0]) < TOL_MEAN
        if n > 1:
            assert relerr(pb[1], pa[1], max(1e-6 * s2, 1e-300)) < TOL_VAR


@pytest.mark.parametrize("n", [131, 136, 200, 250, 391, 520])
def test_ragged_last_tile_against_oracle(kb, n):
    """n % 128 != 0: the kernels skip the identity padding of the last tile (rows of the panel GEMM / solve and of the
    diagonal SYRK, columns of the prediction solve, psi stores).  Likelihood on the left-looking (P = 40) and the
    right-looking (P = 3) schedule and a bulk prediction (many point tiles: the pure-solve kernel + the ragged panel
    kernel) against the oracle; a 3-point-tile batch (per-tile-pair psi CTAs, right-looking sweep from 3 sample tiles
    on) must give the same bits as the same points inside the bulk call."""
    rng = np.random.default_rng(1000 + n)
    d = 4
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X.sum(1, keepdims=True)) + 0.3 * X[:, :1] ** 2 + 3.0
    info = ko.make_kriginfo(X, y)
    pop = rng.uniform(-0.2, 0.5, (40, d))
    nll, inf = kb.likelihood_batch(pop, info, False)
    nll3, inf3 = kb.likelihood_batch(pop[:3], info, False)
    assert not inf.any() and not inf3.any()
    for i in (0, 1, 2, 17, 39):
        ref = ko.likelihood(pop[i].copy(), copy.deepcopy(info), "default", False, check_pd=False)
        assert relerr(nll[i], ref) < TOL_NLL
        if i < 3:
            assert relerr(nll3[i], ref) < TOL_NLL
    A, B = copy.deepcopy(info), copy.deepcopy(info)
    ko.likelihood(pop[0].copy(), A, "all", False, check_pd=False)
    kb.likelihood(pop[0].copy(), B, "all", False)
    assert relerr(B["U"], A["U"]) < TOL_MEAN
    xp = rng.uniform(-1, 1, (30_011, d))
    out = kb.predict_arrays(xp, B, ("pred", "ssqr"))
    sel = np.r_[0:200, 15_000:15_200, 29_811:30_011]
    pa = ko.prediction(xp[sel], A, ["pred", "SSqr"])
    s2 = float(np.asarray(A["SigmaSqr"]).ravel()[0])
    assert relerr(out["pred"][sel], pa[0].ravel()) < TOL_MEAN
    assert relerr(out["ssqr"][sel], pa[1].ravel(), 1e-6 * s2) < TOL_VAR
    sub = kb.predict_arrays(xp[128 * 100:128 * 103], B, ("pred", "ssqr"))
    assert np.array_equal(sub["pred"], out["pred"][128 * 100:128 * 103])
    assert np.array_equal(sub["ssqr"], out["ssqr"][128 * 100:128 * 103])


def test_predict_crosses_chunk_boundary(kb, golden_c3):
    """npred larger than one variance chunk (296 tiles x 128 points at n=200) and not a
    multiple of the tile size; mean-only and mean+variance paths must agree on the mean."""
    g = golden_c3
    info, _ = cases.c3_info(g)
    kb.likelihood(g["fit_x"].copy(), info, "all", False)
    npred = 2 * 296 * 128 + 77
    xp = (ko.mc_points(npred, 6, seed=9) / 2 + 0.5) * (g["ub"] - g["lb"]) + g["lb"]
    f_only = kb.prediction(xp, info, ["pred"])
    f, s = kb.prediction(xp, info, ["pred", "s"])
    assert f_only.shape == (npred, 1) and np.allclose(f_only, f, rtol=1e-12, atol=0)   # two summation orders of psi^T alpha
    sel = np.r_[0:300, 296 * 128 - 150:296 * 128 + 150, npred - 300:npred]
    ra = ko.prediction(xp[sel], info, ["pred", "s"])
    s2 = float(info["SigmaSqr"][0, 0])
    assert relerr(f[sel], ra[0]) < TOL_MEAN
    assert relerr(s[sel], ra[1], 1e-3 * np.sqrt(s2)) < TOL_VAR


def test_host_memory_paths_give_identical_predictions(kb, golden_c3, monkeypatch):
    """Pageable input through the threaded pinned staging ring (h2d_bulk, several trips round the 3 x 32 MiB ring),
    input already pinned, results in recycled pinned arrays or in pageable ones: the same bits every time."""
    from kadal_b200 import _capi
    from kadal_b200.model import model_for
    g = golden_c3
    info, _ = cases.c3_info(g)
    kb.likelihood(g["fit_x"].copy(), info, "all", False)
    m = model_for(info)
    npred = 3_000_011                                       # 144 MB of points
    rng = np.random.default_rng(4)
    xp = rng.uniform(g["lb"], g["ub"], (npred, 6))
    want = 1 | 4
    kw = dict(normtype=1, norm_a=g["lb"], norm_b=g["ub"])
    a = m.predict(xp, want, **kw)
    assert a["pred"].base is not None                        # a view of a pinned buffer
    xpin = _capi.pinned_empty(xp.shape)
    xpin[:] = xp
    b = m.predict(xpin, want, **kw)
    monkeypatch.setattr(_capi, "_RES_LIMIT", 0)
    c = m.predict(xp, want, **kw)
    assert c["pred"].base is None
    for nm in ("pred", "s"):
        assert np.array_equal(a[nm], b[nm]) and np.array_equal(a[nm], c[nm])
    sel = rng.integers(0, npred, 400)
    ra = ko.prediction(xp[sel], info, ["pred", "s"])
    assert relerr(a["pred"][sel], ra[0].ravel()) < TOL_MEAN
    assert relerr(a["s"][sel], ra[1].ravel(), 1e-3 * np.sqrt(float(info["SigmaSqr"][0, 0]))) < TOL_VAR
    _capi.pinned_free(xpin)


def test_return_types_and_errors(kb, golden_c):
    info, x, tv = cases.appendix_c_info(golden_c, "gaussian")
    kb.likelihood(x.copy(), info, "all", tv)
    xp = golden_c["xp"]
    assert isinstance(kb.prediction(xp[0], info, "pred"), float)          # prediction.py:317-321
    assert kb.prediction(xp, info, "SSqr").shape == (3, 1)
    assert kb.prediction(xp, info, "ebe").shape == (1, 3)                 # prediction.py:270
    assert isinstance(kb.prediction(xp, info, ["pred", "s"]), list)
    assert relerr(kb.prediction(xp, info, "FPC"), np.full((3, 1), info["BE"][0, 0])) < 1e-15
    with pytest.raises(NotImplementedError):
        kb.prediction(xp, info, "bogus")
    with pytest.raises(TypeError):
        kb.likelihood(x.copy(), info, "bogus", tv)
    bad = copy.deepcopy(info)
    bad["kernel"] = ["cubic"]
    with pytest.raises(ValueError):
        kb.likelihood(x.copy(), bad, "default", tv)
    with pytest.raises(NotImplementedError):
        kb.prediction(xp, info, "pred", num=0)
    with pytest.raises(ValueError):
        kb.likelihood(np.zeros(9), info, "default", tv)                   # nuggetset: no such layout


def test_empty_population_and_empty_point_set(kb, golden_c):
    info, x, tv = cases.appendix_c_info(golden_c, "gaussian")
    nll, inf = kb.likelihood_batch(np.zeros((0, len(x))), info, tv)
    assert nll.shape == (0,) and inf.shape == (0,)
    kb.likelihood(x.copy(), info, "all", tv)
    res = kb.predict_arrays(np.zeros((0, 4)), info, ("pred", "ssqr"))
    assert res["pred"].shape == (0,) and res["ssqr"].shape == (0,)
    from kadal_b200 import ehvi
    assert ehvi.ehvi3d_sliceupdate_multi(np.zeros((0, 3)), [0.0, 0, 0], np.zeros((0, 3)), np.zeros((0, 3))).shape == (0,)


def test_not_positive_definite_is_reported_per_candidate(kb):
    # KPLS-exponential with signed coefficients gives "distances" of either sign, hence
    # correlations > 1: an indefinite Psi, robustly -- unlike an exactly
    # singular matrix, whose last pivot is rounding noise of either sign.
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (40, 4))
    info = ko.make_kriginfo(X, np.ones((40, 1)) + X[:, :1], pls=rng.standard_normal((4, 2)),
                            n_princomp=2, kernel=("exponential",))
    pop = np.array([[0.0, 0.0], [-0.5, 0.3], [0.5, 0.5]])
    nll, inf = kb.likelihood_batch(pop, info, False)
    assert np.all(inf != 0) and np.all(np.isinf(nll))
    with pytest.raises(np.linalg.LinAlgError):
        kb.likelihood(pop[0], info, "default", False)
    with pytest.raises(np.linalg.LinAlgError):
        ko.likelihood(pop[0], copy.deepcopy(info), "default", False, check_pd=False)
    # healthy candidates of another model are unaffected; a duplicated sample with a vanishing
    # nugget (exactly singular) must give either a flagged candidate or a finite value, never a crash
    good = ko.make_kriginfo(X, np.ones((40, 1)) + X[:, :1])
    n1, i1 = kb.likelihood_batch(rng.uniform(-1, 1, (3, 4)), good, False)
    assert np.all(i1 == 0) and np.all(np.isfinite(n1))
    Xd = X.copy()
    Xd[7] = Xd[3]
    sing = ko.make_kriginfo(Xd, np.ones((40, 1)) + Xd[:, :1], nugget=-300)
    n2, i2 = kb.likelihood_batch(np.zeros((2, 4)), sing, False)
    assert np.all((i2 != 0) == np.isinf(n2)) and not np.any(np.isnan(n2))


def test_debug_naive_kernels_agree_with_dmma(kb):
    from kadal_b200.model import GpuModel
    w = ko.workload("C2", scale=0.3)
    X, y = w["X"], w["y"]
    a = GpuModel(X, y, np.ones((len(X), 1)), [0])
    b = GpuModel(X, y, np.ones((len(X), 1)), [0], naive=True)
    pop = w["popB"][:8]
    na, _ = a.lik_batch(pop, False, -6.0)
    nb, _ = b.lik_batch(pop, False, -6.0)
    assert relerr(na, nb) < 1e-9    # two different factorisation orders (cond up to 5e8)
    a.close()
    b.close()


def test_kpls_and_composite_population(kb):
    rng = np.random.default_rng(8)
    n, d, h = 180, 12, 3
    X = rng.uniform(-1, 1, (n, d))
    y = (X[:, :4].sum(1) ** 2)[:, None] + 3
    pls = rng.standard_normal((d, h)) / 3
    info = ko.make_kriginfo(X, y, pls=pls, n_princomp=h)
    pop = rng.uniform(-1, 1, (12, h))
    ref = ko.likelihood_population(pop, copy.deepcopy(info), False, check_pd=False)
    nll, inf = kb.likelihood_batch(pop, info, False)
    assert relerr(nll, ref) < TOL_NLL and int(np.argmin(nll)) == int(np.argmin(ref))
    info2 = ko.make_kriginfo(X, y, kernel=("gaussian", "matern52"))
    info2["nugget"] = [-8, -2]
    pop2 = np.hstack([rng.uniform(-0.5, 1, (10, d)), rng.uniform(-6, -3, (10, 1)), rng.uniform(0.1, 1, (10, 2))])
    ref2 = ko.likelihood_population(pop2, copy.deepcopy(info2), False, check_pd=False)
    nll2, _ = kb.likelihood_batch(pop2, info2, False)
    assert relerr(nll2, ref2) < TOL_NLL
    popv = np.hstack([pop2[:, :d], rng.uniform(-1, 1, (10, 1)), pop2[:, d:]])
    ref3 = ko.likelihood_population(popv, copy.deepcopy(info2), True, check_pd=False)
    nll3, _ = kb.likelihood_batch(popv, info2, True)
    assert relerr(nll3, ref3) < TOL_NLL


@pytest.mark.parametrize("path", ["full", "left", "right"])
def test_all_factorisation_schedules_against_the_oracle(kb, path, monkeypatch):
    """The library has three batched Cholesky schedules: left-looking and right-looking on half-SM CTAs (two
    per SM, kb200_half.cu; chosen by the parallel work of the call) and the full-SM kernels of the first
    design.  Each is forced here on the same population (n = 700: 6 tiles, ragged last tile) and must meet
    the NLL tolerance, agree on the arg-min, give reproducible bits from run to run and be independent of
    the batch composition."""
    from kadal_b200.model import GpuModel
    monkeypatch.setenv("KB200_HALF", "0" if path == "full" else "1")
    monkeypatch.setenv("KB200_RL", "1" if path == "right" else "0")
    w = ko.workload("C2", scale=0.7)
    n = w["X"].shape[0]
    info = ko.make_kriginfo(w["X"], w["y"])
    pop = np.vstack([w["popB"][:40], w["popA"][:8]])
    gm = GpuModel(w["X"], w["y"], np.ones((n, 1)), [0])
    nll, inf = gm.lik_batch(pop, False, -6.0)
    nll2, _ = gm.lik_batch(pop, False, -6.0)
    assert not inf.any() and np.array_equal(nll, nll2)
    perm = np.random.default_rng(1).permutation(len(pop))
    nll_p, _ = gm.lik_batch(pop[perm], False, -6.0)
    assert np.array_equal(nll_p, nll[perm])
    one, _ = gm.lik_batch(pop[5:6], False, -6.0)
    assert one[0] == nll[5]
    idx = np.r_[np.argsort(nll)[:6], 40, 41]
    ref = ko.likelihood_population(pop[idx], copy.deepcopy(info), False, check_pd=False)
    conds = conds_of(info, pop[idx])
    assert all(relerr(a, b) < nll_tol(c) for a, b, c in zip(nll[idx], ref, conds))
    assert idx[int(np.argmin(ref))] == int(np.argmin(nll))
    st = gm.fit_state(pop[0], False, -6.0)
    A = copy.deepcopy(info)
    ko.likelihood(pop[0].copy(), A, "all", False, check_pd=False)
    Ap = A["Psi"] + np.eye(n) * 1e-6
    assert np.linalg.norm(st["U"].T @ st["U"] - Ap) / np.linalg.norm(Ap) < 1e-14
    gm.close()


def test_results_do_not_depend_on_workspace_chunking_or_stream_groups(kb, monkeypatch):
    """Host-side scheduling must not change a bit: populations larger than the workspace are processed in
    chunks (KB200_WORKSPACE_GB), candidates are split over stream groups (kb200_set_groups) -- every
    candidate is an independent unit."""
    from kadal_b200.model import GpuModel
    w = ko.workload("C2", scale=0.3)
    n = w["X"].shape[0]
    pop = w["popB"][:150]
    gm = GpuModel(w["X"], w["y"], np.ones((n, 1)), [0])
    base, inf = gm.lik_batch(pop, False, -6.0)
    assert not inf.any()
    for g in (1, 4, 8):
        gm.set_groups(g)
        assert np.array_equal(gm.lik_batch(pop, False, -6.0)[0], base)
    gm.close()
    monkeypatch.setenv("KB200_WORKSPACE_GB", "0.01")          # clamped to 256 MiB: ~328 candidates of this size per chunk
    per_mat = 6 * 131072 + 3 * 8192 + 2 * 384 * 8
    gm = GpuModel(w["X"], w["y"], np.ones((n, 1)), [0])
    big = np.vstack([pop, pop[::-1], pop[:77]])                # 377 candidates: two chunks (328 + 49)
    assert len(big) * per_mat > 256 * 2 ** 20
    out, inf = gm.lik_batch(big, False, -6.0)
    assert not inf.any()
    # chunks of one call share the schedule of the call: the repeated candidates give the same bits
    assert np.array_equal(out[:150], out[150:300][::-1]) and np.array_equal(out[:77], out[300:])
    # (the 377-candidate call is left-looking, the 150-candidate one right-looking: equal to rounding)
    assert relerr(out[:150], base) < 1e-9
    gm.close()


def test_kpls_plane_cache_is_bit_identical_to_on_the_fly_assembly(kb, monkeypatch):
    """KPLS distances sum_k g(dx_k) pls_kc do not depend on theta (kernelfunc.py:41-47, 58-64): the
    library caches the h planes per model and assembles every candidate's Psi from them.  Same
    operations in the same order as the on-the-fly kernel -> bit-identical Psi and NLL; both against
    the oracle for gaussian, exponential (signed pls) and a gaussian+exponential composite."""
    from kadal_b200.model import GpuModel
    rng = np.random.default_rng(18)
    n, d, h = 300, 14, 3
    X = rng.uniform(-1, 1, (n, d))
    y = (X[:, :4].sum(1) ** 2)[:, None] + 3
    pls = np.abs(rng.standard_normal((d, h))) / 3     # positive: the exponential KPLS kernel stays PD
    pop = rng.uniform(-0.3, 0.8, (6, h))
    for kernels, extra in ((("gaussian",), 0), (("exponential",), 0), (("gaussian", "exponential"), 2)):
        info = ko.make_kriginfo(X, y, kernel=kernels, pls=pls, n_princomp=h)
        pp = np.hstack([pop, rng.uniform(0.2, 1, (6, extra))]) if extra else pop
        ref = ko.likelihood_population(pp, copy.deepcopy(info), False, check_pd=False)
        kids = [{"gaussian": 0, "exponential": 1}[k] for k in kernels]
        res = {}
        for on in ("1", "0"):
            monkeypatch.setenv("KB200_KPLS_PLANES", on)
            gm = GpuModel(X, y, np.ones((n, 1)), kids, pls=pls)
            nll, inf = gm.lik_batch(pp, False, -6.0)
            psi = gm.debug_psi(pp[0], False, -6.0)
            res[on] = (nll, psi)
            assert not inf.any()
            gm.close()
        assert np.array_equal(res["1"][0], res["0"][0]) and np.array_equal(res["1"][1], res["0"][1])
        A = copy.deepcopy(info)
        ko.likelihood(pp[0].copy(), A, "all", False, check_pd=False)
        assert np.max(np.abs(res["1"][1] - A["Psi"])) < 1e-14
        conds = conds_of(info, pp)
        assert all(relerr(a, b) < nll_tol(c) for a, b, c in zip(res["1"][0], ref, conds))


# ------------------------------------------------------------------ drop-in classes (BASELINE config 1)
def _branin(X):
    a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6.0, 10.0, 1 / (8 * np.pi)
    x1, x2 = X[:, 0], X[:, 1]
    return (a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s)[:, None]


def test_krigdemo_train_and_predict_drop_in(kb):
    """KrigDemo-equivalent (KrigDemo.py:22-118): Ordinary Kriging, Gaussian kernel, Branin 2-D,
    50 low-discrepancy samples, L-BFGS-B with 7 restarts, predict a 100x100 grid.  The trained
    KrigInfo is then handed to the oracle: same NLL at the optimum, same predictions."""
    from kadal_b200 import Kriging
    from scipy.stats import qmc
    lb, ub = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    X = qmc.scale(qmc.Sobol(2, scramble=False).random(51)[1:], lb, ub)
    y = _branin(X)
    K = {"X": X, "y": y, "lb": lb, "ub": ub, "nvar": 2, "problem": "branin", "nugget": -6,
         "nrestart": 7, "kernel": ["gaussian"], "optimizer": "lbfgsb"}
    model = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    model.train(disp=False)
    info = model.KrigInfo
    assert model.train_stats["evals"] > model.train_stats["batches"] >= 1     # restarts ran in lock step
    # the optimum the GPU search found is a genuine improvement over the start points' box centre
    ref = copy.deepcopy(info)
    nll_ref = ko.likelihood(np.array(info["Theta"], dtype=float), ref, "all", False, check_pd=False)["NegLnLike"]
    n = X.shape[0]
    cond = np.linalg.cond(ref["Psi"] + np.eye(n) * 1e-6)
    assert relerr(info["NegLnLike"], nll_ref) < nll_tol(cond)
    centre = ko.likelihood(np.zeros(2), copy.deepcopy(info), "default", False, check_pd=False)
    assert info["NegLnLike"] <= centre + 1e-9
    g1, g2 = np.meshgrid(np.linspace(lb[0], ub[0], 100), np.linspace(lb[1], ub[1], 100))
    xg = np.column_stack([g1.ravel(), g2.ravel()])
    p = model.predict(xg, ["pred", "SSqr"])
    pr = ko.prediction(xg, ref, ["pred", "SSqr"])
    s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    scale = float(np.max(np.abs(y)))
    assert np.max(np.abs(p[0] - pr[0])) / scale < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    assert relerr(p[1].ravel(), pr[1].ravel(), 1e-6 * s2) < max(TOL_VAR, 5 * cond * 2.0 ** -53)
    # the surrogate interpolates the samples (nugget 1e-6)
    assert np.max(np.abs(model.predict(X, ["pred"]) - y)) < 1e-3 * scale


def test_trained_model_round_trips_through_pickle_and_believer_refit(kb, tmp_path):
    """KADAL pickles trained Kriging objects (mobo_3d.py:244-248) and refits at
    fixed theta after appending a sample (Kriging believer, MOBO.py:344-356; AK-MCS / SOBO updates).  The
    drop-in keeps KrigInfo a plain dict and the device state outside it: a pickled / unpickled model
    predicts the same values, and the believer refit with one more sample matches the oracle."""
    import pickle
    from kadal_b200 import Kriging
    from kadal_b200.model import clear_cache
    from kadal_b200.trendfunction import compute_regression_mat
    rng = np.random.default_rng(23)
    n, d = 60, 3
    lb, ub = -np.ones(d), np.ones(d)
    X = rng.uniform(-1, 1, (n, d))
    y = (np.sin(2 * X[:, 0]) + X[:, 1] * X[:, 2])[:, None] + 2.0
    K = {"X": X, "y": y, "lb": lb, "ub": ub, "nvar": d, "nugget": -6, "nrestart": 3, "kernel": ["gaussian"],
         "optimizer": "lbfgsb"}
    model = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    model.train(disp=False)
    xq = rng.uniform(-1, 1, (200, d))
    before = model.predict(xq, ["pred", "s"])
    blob = tmp_path / "krig.pkl"
    with open(blob, "wb") as f:
        pickle.dump(model, f)
    clear_cache()                                   # as in a fresh process: no device state survives
    with open(blob, "rb") as f:
        again = pickle.load(f)
    after = again.predict(xq, ["pred", "s"])
    # (the reloaded predictor is rebuilt from KrigInfo['U'] / 'BE' instead of the device-resident fit:
    #  same numbers to rounding, not the same bits)
    assert relerr(after[0], before[0]) < 1e-10 and relerr(after[1], before[1], 1e-3 * float(np.max(before[1]))) < 1e-8
    # Kriging believer: append the prediction at xnext as a pseudo-observation, refit at the same theta
    xnext = rng.uniform(-1, 1, (1, d))
    ynext = again.predict(xnext, ["pred"])
    s_before = float(again.predict(xnext, ["s"]))
    info = again.KrigInfo
    info["X"] = np.vstack((info["X"], xnext))
    info["y"] = np.vstack((info["y"], np.reshape(ynext, (1, 1))))
    info["nsamp"] = n + 1
    again.standardize()
    info = again.KrigInfo
    info["F"] = compute_regression_mat(info["idx"], info["X_norm"], np.vstack((-np.ones((1, d)), np.ones((1, d)))), np.ones(d))
    ref = copy.deepcopy(info)
    again.KrigInfo = kb.likelihood(info["Theta"], info, mode="all", trainvar=False)
    ko.likelihood(np.array(ref["Theta"], dtype=float), ref, "all", False, check_pd=False)
    cond = np.linalg.cond(ref["Psi"] + np.eye(n + 1) * 1e-6)
    assert relerr(again.KrigInfo["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
    p2, r2 = again.predict(xq, ["pred", "s"]), ko.prediction(xq, ref, ["pred", "s"])
    assert relerr(p2[0], r2[0]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    # the believer point is reproduced and carries (almost) no variance
    assert abs(float(again.predict(xnext, ["pred"])) - float(ynext)) < 1e-6
    s_after, s_ref = float(again.predict(xnext, ["s"])), float(ko.prediction(xnext, ref, ["s"]))
    assert s_after < s_before and abs(s_after - s_ref) < max(TOL_VAR, 5 * cond * 2.0 ** -53) * max(s_ref, 1e-3)


def _believer_step(kb, info, xnew_norm, ynew):
    """What MOBO.py:344-356 does to a KrigInfo between two likelihood(..., 'all') calls (standardisation off here:
    X_norm of the old rows does not change under 'default' standardisation either)."""
    info["X"] = np.vstack((info["X"], xnew_norm))
    info["X_norm"] = info["X"]
    info["y"] = np.vstack((info["y"], np.reshape(ynew, (1, 1))))
    info["nsamp"] = info["X"].shape[0]
    info["F"] = np.ones((info["nsamp"], 1))
    return info


@pytest.mark.parametrize("kernel,n0", [(["gaussian"], 125), (["matern52"], 60), (["gaussian", "exponential"], 254)])
def test_one_sample_append_matches_refactorisation_and_oracle(kb, monkeypatch, kernel, n0):
    """likelihood(Theta, KrigInfo, 'all') after one stacked sample extends the cached
    factor by a row (kb200_fit_append, O(n^2)) -- six appends in a row, across a 128-tile boundary, each checked
    against the oracle and against the full refactorisation ($KB200_NO_APPEND=1) of the same KrigInfo."""
    from kadal_b200.model import model_for, clear_cache
    clear_cache()
    rng = np.random.default_rng(31)
    d = 4
    X = rng.uniform(-1, 1, (n0, d))
    f = lambda Z: np.sin(Z.sum(1)) + 0.3 * Z[:, 0] ** 2 + 3.0
    info = ko.make_kriginfo(X, f(X)[:, None], kernel=kernel, nugget=-6)
    x = np.array([0.1, -0.2, 0.3, 0.0])
    if len(kernel) > 1:
        x = np.concatenate((x, [0.7, 0.4]))           # kernel weights ride along in x (nuggetset)
    kb.likelihood(x.copy(), info, "all", False)
    assert model_for(info).last_fit_appended is False
    xq = rng.uniform(-1, 1, (300, d))
    for step in range(6):
        xn = rng.uniform(-1, 1, (1, d))
        yn = kb.prediction(xn, info, ["pred"]) if step % 2 == 0 else f(xn)     # believer / true observation
        info = _believer_step(kb, info, xn, yn)
        ref, full = copy.deepcopy(info), copy.deepcopy(info)
        kb.likelihood(x.copy(), info, "all", False)
        assert model_for(info).last_fit_appended is True
        n = info["X"].shape[0]
        ko.likelihood(x.copy(), ref, "all", False, check_pd=False)
        cond = np.linalg.cond(ref["Psi"] + np.eye(n) * 1e-6)
        assert np.array_equal(info["Psi"], ref["Psi"]) or relerr(info["Psi"], ref["Psi"]) < 1e-15
        assert relerr(info["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
        assert relerr(info["BE"], ref["BE"]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
        assert relerr(info["SigmaSqr"], ref["SigmaSqr"]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
        assert relerr(info["U"], ref["U"], 1.0) < max(1e-10, 0.5 * cond * 2.0 ** -53)
        assert np.all(np.tril(info["U"], -1) == 0.0)
        got, want = kb.prediction(xq, info, ["pred", "SSqr"]), ko.prediction(xq, ref, ["pred", "SSqr"])
        assert relerr(got[0], want[0]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
        s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
        assert relerr(got[1], want[1], 1e-6 * s2) < max(TOL_VAR, 5 * cond * 2.0 ** -53)
        # the same KrigInfo refactorised from scratch on the device
        monkeypatch.setenv("KB200_NO_APPEND", "1")
        clear_cache_keep = model_for(full)            # a second model for the copied arrays
        kb.likelihood(x.copy(), full, "all", False)
        assert clear_cache_keep.last_fit_appended is False
        monkeypatch.delenv("KB200_NO_APPEND")
        assert relerr(info["NegLnLike"], full["NegLnLike"]) < max(1e-11, 0.05 * cond * 2.0 ** -53)
        assert relerr(info["U"], full["U"], 1.0) < max(1e-12, 0.05 * cond * 2.0 ** -53)
    clear_cache()


def test_one_sample_append_kpls_and_trainvar(kb):
    """The append under KPLS-projected distances (the theta-independent planes of the bordered model are rebuilt, the
    factor is extended) and with trainvar=True (sigma^2 rides in x and does not enter Psi)."""
    from kadal_b200.model import model_for, clear_cache
    clear_cache()
    rng = np.random.default_rng(77)
    n0, d, h = 127, 6, 2
    X = rng.uniform(-1, 1, (n0, d))
    f = lambda Z: (Z[:, :3].sum(1)) ** 2 + 0.5 * Z[:, 4] + 3.0
    pls = rng.standard_normal((d, h)) / 3.0
    for tv, x in ((False, np.array([0.2, -0.1])), (True, np.array([0.2, -0.1, 0.3]))):
        info = ko.make_kriginfo(X, f(X)[:, None], kernel=["gaussian"], nugget=-6, pls=pls, n_princomp=h)
        kb.likelihood(x.copy(), info, "all", tv)
        for step in range(3):
            xn = rng.uniform(-1, 1, (1, d))
            info = _believer_step(kb, info, xn, f(xn))
            ref = copy.deepcopy(info)
            kb.likelihood(x.copy(), info, "all", tv)
            assert model_for(info).last_fit_appended is True
            ko.likelihood(x.copy(), ref, "all", tv, check_pd=False)
            n = info["X"].shape[0]
            cond = np.linalg.cond(ref["Psi"] + np.eye(n) * 1e-6)
            assert relerr(info["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
            assert relerr(info["U"], ref["U"], 1.0) < max(1e-10, 0.5 * cond * 2.0 ** -53)
            xq = rng.uniform(-1, 1, (100, d))
            assert relerr(kb.prediction(xq, info, ["pred"]), ko.prediction(xq, ref, ["pred"])) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    clear_cache()


def test_append_refuses_other_hyperparameters_and_falls_back_to_full_fit(kb):
    """kb200_fit_append fails loudly when the hyper-parameters differ; the host layer only appends when the cached
    fit is the fit of the same x (otherwise likelihood(..., 'all') refactorises on the device)."""
    from kadal_b200 import _capi
    from kadal_b200.model import model_for, clear_cache
    clear_cache()
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (40, 3))
    info = ko.make_kriginfo(X, np.sin(X.sum(1))[:, None] + 2.0, kernel=["gaussian"], nugget=-6)
    x = np.array([0.0, 0.1, -0.1])
    kb.likelihood(x.copy(), info, "all", False)
    prev = model_for(info)
    info = _believer_step(kb, info, rng.uniform(-1, 1, (1, 3)), np.array([2.0]))
    nxt = model_for(info)
    with pytest.raises(_capi.Kb200Error):
        nxt.fit_state(x + 0.01, False, -6.0, append_to=prev)
    ref = copy.deepcopy(info)
    kb.likelihood(x + 0.01, info, "all", False)                 # other theta: full fit, no append
    assert nxt.last_fit_appended is False
    ko.likelihood(x + 0.01, ref, "all", False, check_pd=False)
    assert relerr(info["NegLnLike"], ref["NegLnLike"]) < 1e-9
    clear_cache()


def test_kpls_class_cmaes_drop_in(kb):
    """KPLS (kpls_model.py): PLS rotations on the host (scikit-learn, as the reference), CMA-ES
    populations evaluated in GPU batches, prediction through the projected distances."""
    from kadal_b200 import KPLS
    rng = np.random.default_rng(3)
    n, d = 150, 12
    X = rng.uniform(-1, 1, (n, d))
    y = (X[:, :4].sum(1) ** 2)[:, None] + 0.5 * X[:, 5:6] + 3.0
    K = {"X": X, "y": y, "lb": -np.ones(d), "ub": np.ones(d), "nvar": d, "nugget": -6, "n_princomp": 3,
         "nrestart": 2, "kernel": ["gaussian"], "optimizer": "cmaes"}
    model = KPLS(K, standardization=True, standtype="default", normy=False, trainvar=False)
    model.train(disp=False, seed=1)
    info = model.KrigInfo
    assert info["plscoeff"].shape == (d, 3) and info["type"] == "kpls"
    ref = copy.deepcopy(info)
    ko.likelihood(np.array(info["Theta"], dtype=float), ref, "all", False, check_pd=False)
    cond = np.linalg.cond(ref["Psi"] + np.eye(n) * 1e-6)
    assert relerr(info["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
    xp = rng.uniform(-1, 1, (500, d))
    p = model.predict(xp, ["pred", "SSqr"])
    pr = ko.prediction(xp, ref, ["pred", "SSqr"])
    s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    assert relerr(p[0].ravel(), pr[0].ravel()) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    assert relerr(p[1].ravel(), pr[1].ravel(), 1e-6 * s2) < max(TOL_VAR, 5 * cond * 2.0 ** -53)


# ------------------------------------------------------------------ BASELINE configs 3, 4, 5 at full size
def test_c3_full_size_bulk_predict_properties(kb):
    """C3: n=200, d=6, pred + s + U over 10^7 Monte-Carlo points in ONE call (the reference chunks
    to 10 000 rows, akmcs.py:91-106).  Size-independent properties: any slice of the bulk result
    equals the same points predicted on their own (bit-identical), a random sample of 4000 points
    matches the oracle, U = |f|/s, and the fused small-n kernel agrees with the general tiled path."""
    w = ko.workload("C3")
    info = ko.make_kriginfo(w["X"], w["y"])
    kb.likelihood(w["theta"].copy(), info, "all", False)
    ref = copy.deepcopy(info)
    npts = 10 ** 7
    xp = ko.mc_points(npts, 6, seed=2)
    out = kb.predict_arrays(xp, info, ("pred", "s", "ufun"))
    f, s, u = out["pred"], out["s"], out["ufun"]
    assert f.shape == (npts,) and np.all(np.isfinite(f)) and np.all(s >= 0)
    assert np.array_equal(u, np.abs(f) / s)
    sl = slice(4_999_871, 5_000_129)                  # straddles point-tile and chunk boundaries
    sub = kb.predict_arrays(xp[sl], info, ("pred", "s"))
    assert np.array_equal(sub["pred"], f[sl]) and np.array_equal(sub["s"], s[sl])
    idx = np.random.default_rng(4).choice(npts, 4000, replace=False)
    pr = ko.prediction(xp[idx], ref, ["pred", "s"])
    s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    assert relerr(f[idx], pr[0].ravel()) < TOL_MEAN
    assert relerr(s[idx], pr[1].ravel(), 1e-3 * np.sqrt(s2)) < TOL_VAR
    # argmin of the U-function (the AK-MCS learning criterion, akmcs.py:281-285) on the sample
    assert int(np.argmin(u[idx])) == int(np.argmin(np.abs(pr[0].ravel()) / pr[1].ravel()))
    # the fused small-n kernel against the general tiled path (debug kernels: plain-FMA GEMM, per-row
    # substitution) on the same model
    from kadal_b200 import _capi
    from kadal_b200.model import GpuModel
    gm = GpuModel(w["X"], w["y"], np.ones((200, 1)), [0], naive=True)
    assert gm.fit_state(w["theta"], False, -6.0, want_U=False, want_Psi=False)["info"] == 0
    o2 = gm.predict(xp[:5000], _capi.OUT_PRED | _capi.OUT_S, _capi.NORM_DEFAULT, -np.ones(6), np.ones(6))
    assert relerr(o2["pred"], f[:5000]) < 1e-11
    assert relerr(o2["s"], s[:5000], 1e-3 * np.sqrt(s2)) < 1e-8
    gm.close()


def test_c4_full_model_prediction_against_oracle(kb):
    """C4: n=800, d=20 (7 sample tiles), 'pred' only as sobol_ind.py:45-71 requests; 2x10^5 points
    here (the 2x10^7 of the config are the same kernel, sharded), oracle on a 3000-point sample."""
    w = ko.workload("C4")
    info = ko.make_kriginfo(w["X"], w["y"])
    kb.likelihood(w["theta"].copy(), info, "all", False)
    ref = copy.deepcopy(info)
    ko.likelihood(w["theta"].copy(), ref, "all", False, check_pd=False)
    cond = np.linalg.cond(ref["Psi"] + np.eye(800) * 1e-6)
    assert relerr(info["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
    xp = np.random.default_rng(12).uniform(-1, 1, (200_000, 20))
    f = kb.predict_arrays(xp, info, ("pred",))["pred"]
    idx = np.random.default_rng(13).choice(len(xp), 3000, replace=False)
    pr = ko.prediction(xp[idx], ref, ["pred"])
    assert relerr(f[idx], pr.ravel()) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    # mean + variance through the general tiled path (n > 256) on a subset
    pv = kb.predict_arrays(xp[idx[:600]], info, ("pred", "ssqr"))
    rv = ko.prediction(xp[idx[:600]], ref, ["pred", "SSqr"])
    s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    assert relerr(pv["pred"], rv[0].ravel()) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    assert relerr(pv["ssqr"], rv[1].ravel(), 1e-6 * s2) < max(TOL_VAR, 5 * cond * 2.0 ** -53)


def test_small_batches_on_a_large_model_take_the_right_looking_sweep(kb, monkeypatch):
    """Candidate scoring predicts a few hundred points on a model of many sample tiles (C4: nt = 7): one CTA per (point tile,
    sample tile) pair for psi and a right-looking solve / update sweep instead of nt sequential left-looking launches on a
    handful of SMs.  Both sweeps against the oracle, and against each other."""
    w = ko.workload("C4")
    info = ko.make_kriginfo(w["X"], w["y"])
    kb.likelihood(w["theta"].copy(), info, "all", False)
    ref = copy.deepcopy(info)
    ko.likelihood(w["theta"].copy(), ref, "all", False, check_pd=False)
    cond = np.linalg.cond(ref["Psi"] + np.eye(800) * 1e-6)
    s2 = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    xp = np.random.default_rng(99).uniform(-1, 1, (777, 20))
    rv = ko.prediction(xp, ref, ["pred", "SSqr"])
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KB200_PRED_RL", mode)
        got[mode] = kb.predict_arrays(xp, info, ("pred", "ssqr"))
        assert relerr(got[mode]["pred"], rv[0].ravel()) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
        assert relerr(got[mode]["ssqr"], rv[1].ravel(), 1e-6 * s2) < max(TOL_VAR, 5 * cond * 2.0 ** -53)
    monkeypatch.delenv("KB200_PRED_RL")
    # same products accumulated from -psi in the same order k = 0, 1, ... (in registers, or tile by tile through memory):
    # the two sweeps give the same bits, so a point's variance does not depend on the size of the batch it came in
    assert np.array_equal(got["1"]["ssqr"], got["0"]["ssqr"])
    assert relerr(got["1"]["pred"], got["0"]["pred"]) < 1e-13      # psi^T alpha: per-sample-tile partial sums vs one running sum
    auto = kb.predict_arrays(xp, info, ("pred", "ssqr"))                       # 7 point tiles: right-looking by default
    assert np.array_equal(auto["ssqr"], got["1"]["ssqr"]) and np.array_equal(auto["pred"], got["1"]["pred"])
    one = kb.predict_arrays(xp[:1], info, ("pred",))["pred"]                   # mean only, a single point: split psi kernel
    assert relerr(one, rv[0].ravel()[:1]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)


def test_c5_kpls_n4000_factor_and_candidate_scoring(kb):
    """C5: KPLS, d=100, n=4000 (32 tiles per dimension: 528 factor tiles, blocked tensor-core
    Cholesky), 3 PLS components.  The oracle needs a 12.8 GB temporary and ~1 min per evaluation at
    this size, so full size is checked through properties: U^T U = Psi + eps I, interpolation of the
    samples, positive variances, batched == single evaluation; the oracle is used at n=500."""
    w = ko.workload("C5")
    info = ko.make_kriginfo(w["X"], w["y"], pls=w["pls"], n_princomp=3)
    x = w["theta"]
    kb.likelihood(x.copy(), info, "all", False)
    U, Psi = info["U"], info["Psi"]
    n = 4000
    A = Psi + np.eye(n) * 1e-6
    assert np.allclose(np.tril(U, -1), 0.0)
    assert np.linalg.norm(U.T @ U - A) / np.linalg.norm(A) < 1e-14
    # Psi itself against the kernel definition on a 300 x 300 corner (kernelfunc.py:42-47)
    Xn = info["X_norm"] if "X_norm" in info else info["X"]
    dd = (Xn[:300, None, :] - Xn[None, :300, :]) ** 2
    S = ((dd @ (w["pls"] ** 2)) / (10.0 ** x) ** 2).sum(-1)
    assert np.max(np.abs(np.exp(-0.5 * S) - Psi[:300, :300])) < 1e-14
    nll_b, inf = kb.likelihood_batch(np.vstack([x, x + 0.05]), info, False)
    assert inf[0] == 0 and relerr(nll_b[0], info["NegLnLike"]) < 1e-13
    # NLL / BE / sigma^2 at full size against a lean float64 Cholesky (LAPACK dpotrf + two triangular solves) of the
    # exported Psi -- SURVEY.md Appendix A steps 3-5, which the survey measured equal to the reference's six general
    # solves to <= 1.1e-13 -- with the bound scaled by cond(Psi + eps I) as everywhere else
    from scipy.linalg import solve_triangular
    Lc = np.linalg.cholesky(A)
    a = solve_triangular(Lc, info["y"], lower=True)
    b = solve_triangular(Lc, np.ones((n, 1)), lower=True)
    beta = float((b.T @ a) / (b.T @ b))
    r = a - b * beta
    s2_ref = float(r.T @ r) / n
    nll_ref = 0.5 * n * np.log(s2_ref) + np.sum(np.log(np.diag(Lc)))
    ev = np.linalg.eigvalsh(A)
    cond = ev[-1] / ev[0]
    assert relerr(info["NegLnLike"], nll_ref) < nll_tol(cond), (info["NegLnLike"], nll_ref, cond)
    assert relerr(info["BE"][0, 0], beta) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    assert relerr(float(info["SigmaSqr"][0, 0]), s2_ref, 1e-300) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    # EHVI-style candidate scoring: pred + SSqr for a [n_pop, d] population (EHVIcomputation.py:169-172)
    cand = np.random.default_rng(21).uniform(-1, 1, (1500, 100))
    p = kb.predict_arrays(np.vstack([info["X"][:200], cand]), info, ("pred", "ssqr"))
    s2 = float(np.asarray(info["SigmaSqr"]).ravel()[0])
    assert np.max(np.abs(p["pred"][:200] - info["y"][:200, 0])) < 1e-3 * np.ptp(info["y"])
    assert np.all(p["ssqr"][200:] > -1e-9 * s2) and np.all(p["ssqr"][:200] < 1e-3 * s2)
    # oracle parity of the same model family at a size the oracle finishes in seconds
    ws = ko.workload("C5", scale=0.125)
    si = ko.make_kriginfo(ws["X"], ws["y"], pls=ws["pls"], n_princomp=3)
    ref = copy.deepcopy(si)
    kb.likelihood(x.copy(), si, "all", False)
    ko.likelihood(x.copy(), ref, "all", False, check_pd=False)
    cond = np.linalg.cond(ref["Psi"] + np.eye(500) * 1e-6)
    assert relerr(si["NegLnLike"], ref["NegLnLike"]) < nll_tol(cond)
    q = kb.prediction(cand[:400], si, ["pred", "SSqr"])
    qr = ko.prediction(cand[:400], ref, ["pred", "SSqr"])
    s2s = float(np.asarray(ref["SigmaSqr"]).ravel()[0])
    assert relerr(q[0], qr[0]) < max(TOL_MEAN, 0.5 * cond * 2.0 ** -53)
    assert relerr(q[1], qr[1], 1e-6 * s2s) < max(TOL_VAR, 5 * cond * 2.0 ** -53)


def test_prediction_exp_is_libm_exp_without_the_denormal_path(kb):
    """kb200_corr.cuh:exp_nonpos (what the prediction kernels call: no branch, constants from the constant bank) against
    libm's exp on the same device: bit-identical on [-708, 0], clamped below; and against numpy's exp to 1 ulp."""
    from kadal_b200 import _capi
    lib = _capi.require_gpu()
    rng = np.random.default_rng(6)
    x = np.concatenate([-rng.uniform(0, 708, 200_000), -10.0 ** rng.uniform(-300, 2.8, 100_000), [0.0, -708.0, -1e-320, -0.0],
                        -rng.uniform(708, 1e12, 1000), [-np.inf]])
    x = np.ascontiguousarray(x)
    mine, ref = np.empty_like(x), np.empty_like(x)
    _capi.check(lib.kb200_debug_exp(0, _capi.ptr(x), len(x), 0, _capi.ptr(mine)))
    _capi.check(lib.kb200_debug_exp(0, _capi.ptr(x), len(x), 1, _capi.ptr(ref)))
    inside = x >= -708.0
    assert np.array_equal(mine[inside], ref[inside])
    at_clamp = mine[np.flatnonzero(x == -708.0)[0]]
    assert 3.3e-308 < at_clamp < 3.4e-308 and np.all(mine[~inside] == at_clamp) and np.all(ref[~inside] < 3.4e-308)
    npx = np.exp(x[inside])
    assert np.max(np.abs(mine[inside] - npx) / npx) < 2.3e-16


# ------------------------------------------------------------------ small n: one CTA per matrix in shared memory
@pytest.mark.parametrize("n,d,kernel,tv", [(50, 2, ("gaussian",), False), (64, 4, ("matern52",), False),
                                           (190, 6, ("gaussian",), False), (160, 6, ("gaussian",), True),
                                           (97, 10, ("gaussian", "exponential"), False), (192, 3, ("exponential",), False),
                                           (33, 1, ("gaussian",), False)])
def test_small_n_in_shared_memory_likelihood_against_oracle_and_tile_path(kb, monkeypatch, n, d, kernel, tv):
    """lik_small_kernel (north star (2): one CTA per matrix held in shared memory -- assembly, blocked POTF2 with warp
    shuffles, fused forward substitution; n <= 192, 256 threads and two CTAs per SM up to n = 128, 512 threads beyond) against the oracle at the north-star tolerance
    and against the tile path ($KB200_NO_SMALL_LIK=1) it replaces for these sizes: same Psi bits, another order of the
    factorisation's roundings."""
    rng = np.random.default_rng(1000 * n + d)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X.sum(1, keepdims=True)) + 0.2 * X[:, :1] ** 2 + 3.0
    info = ko.make_kriginfo(X, y, kernel=kernel)
    nk = len(kernel)
    P = 40
    pop = rng.uniform(-1.0, 1.0, (P, d))
    if tv:
        pop = np.hstack([pop, rng.uniform(-1, 1, (P, 1))])
    if nk > 1:
        pop = np.hstack([pop, rng.uniform(0.1, 1.0, (P, nk))])
    from kadal_b200.model import model_for
    l0 = model_for(info).launch_count()
    nll, inf = kb.likelihood_batch(pop, info, tv)
    assert model_for(info).launch_count() - l0 == 3          # decode, lik_small_kernel, lik_finalize_kernel
    assert np.all(inf == 0)
    ref = ko.likelihood_population(pop, copy.deepcopy(info), tv, check_pd=False)
    conds = conds_of(info, pop) if not tv else conds_of(info, np.delete(pop, d, 1))
    rel = np.abs(nll - ref) / np.maximum(np.abs(ref), 1.0)
    assert np.all(rel < np.array([nll_tol(c) for c in conds])), (rel / np.array([nll_tol(c) for c in conds])).max()
    assert int(np.argmin(nll)) == int(np.argmin(ref))
    # batch composition does not matter: one CTA per candidate, no cross-talk
    perm = rng.permutation(P)
    assert np.array_equal(kb.likelihood_batch(pop[perm], info, tv)[0], nll[perm])
    assert kb.likelihood_batch(pop[7:8], info, tv)[0][0] == nll[7]
    monkeypatch.setenv("KB200_NO_SMALL_LIK", "1")
    l0 = model_for(info).launch_count()
    tile, inf2 = kb.likelihood_batch(pop, info, tv)
    assert model_for(info).launch_count() - l0 > 3           # assembly, diagonal / panel tiles, finalize
    monkeypatch.delenv("KB200_NO_SMALL_LIK")
    assert np.all(inf2 == 0)
    relt = np.abs(nll - tile) / np.maximum(np.abs(tile), 1.0)
    assert np.all(relt < np.array([nll_tol(c) for c in conds]))
    if n <= 128:      # one tile: the same blocked POTF2 on the same Psi bits -> the same result, bit for bit
        assert np.array_equal(nll, tile)


def test_small_n_path_reports_non_positive_definite_candidates(kb):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (40, 4))
    info = ko.make_kriginfo(X, np.ones((40, 1)) + X[:, :1], pls=rng.standard_normal((4, 2)), n_princomp=2,
                            kernel=("exponential",))
    good = ko.make_kriginfo(X, np.ones((40, 1)) + X[:, :1])
    nll, inf = kb.likelihood_batch(np.array([[0.0, 0.0], [-0.5, 0.3]]), info, False)
    assert np.all(inf > 0) and np.all(inf <= 40) and np.all(np.isinf(nll))
    nll, inf = kb.likelihood_batch(np.zeros((3, 4)), good, False)
    assert np.all(inf == 0) and np.all(np.isfinite(nll))


# ------------------------------------------------------------------ LOOCV
def _loocv2_bruteforce(info):
    """krigloocv2.py:7-34 restated with the oracle: n refits at fixed theta, one-point predictions."""
    n = info["nsamp"]
    out = np.zeros((n, 1))
    for i in range(n):
        t = copy.deepcopy(info)
        t["F"] = np.delete(info["F"], i, 0)
        t["y"] = np.delete(info["y"], i, 0)
        t["X_norm"] = np.delete(info["X_norm"], i, 0)
        t["nsamp"] = n - 1
        t = ko.likelihood(np.array(info["Theta"], dtype=float), t, "all", False, check_pd=False)
        t["standardization"] = False
        t["X"] = t["X_norm"]
        out[i, 0] = ko.prediction(info["X_norm"][i, :], t, ["pred"])
    return out


@pytest.mark.parametrize("n,d,kernel", [(70, 3, "gaussian"), (150, 4, "matern52"), (260, 5, "gaussian")])
def test_loocv_standard_matches_n_refits(kb, n, d, kernel):
    from kadal_b200.loocv import loocv
    from kadal_b200.errperf import errperf
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(2 * X.sum(1, keepdims=True)) + X[:, :1] ** 2 + 3.0
    info = ko.make_kriginfo(X, y, kernel=(kernel,))
    kb.likelihood(np.full(d, -0.1), info, "all", False)
    err, pred = loocv(info, errtype="rmse", type="standard")
    ref = _loocv2_bruteforce(info)
    cond = np.linalg.cond(info["Psi"] + np.eye(n) * 1e-6)
    assert pred.shape == (n, 1)
    assert relerr(pred, ref) < max(TOL_MEAN, 5 * cond * 2.0 ** -53)
    assert abs(err - errperf(y, ref, "rmse")) <= 1e-7 * max(1.0, abs(err))
    with pytest.raises(ValueError):
        loocv(info, type="other")


def _loocv_fast_restated(info):
    """krigloocv.py:4-38 as written: inverse of [[sigma^2 Psi, F], [F^T, 1]] (a ONE in the corner, Psi without nugget)."""
    n = info["nsamp"]
    B = np.linalg.inv(np.block([[info["SigmaSqr"] * info["Psi"], info["F"]], [info["F"].T, np.ones((1, 1))]]))
    y = info["y"][:, 0]
    return (y - (B[:n, :n] @ y) / np.diag(B)[:n]).reshape(-1, 1)


@pytest.mark.parametrize("n,d,kernel,lt", [(70, 3, "gaussian", -0.3), (150, 4, "matern52", 0.3), (260, 5, "gaussian", 0.0),
                                           (90, 2, "exponential", 0.0)])
def test_loocv_fast_matches_the_bordered_inverse(kb, n, d, kernel, lt):
    """type='fast' (krigloocv.py): closed form over the factor of Psi WITHOUT nugget against the reference's dense inverse."""
    from kadal_b200.loocv import loocv
    from kadal_b200.errperf import errperf
    rng = np.random.default_rng(n + 1)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(2 * X.sum(1, keepdims=True)) + X[:, :1] ** 2 + 3.0
    info = ko.make_kriginfo(X, y, kernel=(kernel,))
    kb.likelihood(np.full(d, lt), info, "all", False)
    err, pred = loocv(info, errtype="rmse", type="fast")
    ref = _loocv_fast_restated(info)
    cond = np.linalg.cond(info["Psi"])
    assert pred.shape == (n, 1)
    assert relerr(pred, ref) < max(TOL_MEAN, 5 * cond * 2.0 ** -53)
    assert abs(err - errperf(y, ref, "rmse")) <= 1e-7 * max(1.0, abs(err))


# ------------------------------------------------------------------ AK-MCS reductions
def test_akmcs_reductions_match_the_reference_formulas(kb):
    """lfucalc / pfcalc / stopcrit / covpf of akmcs.py:267-297 on a 1.3e6-point population: the
    device-side reduction must equal the same formulas applied to the bulk predictions."""
    from kadal_b200.reliability import akmcs_reductions
    w = ko.workload("C3")
    y = w["y"] - 3.6                                   # limit state G = 0 cuts through the population
    info = ko.make_kriginfo(w["X"], y)
    kb.likelihood(w["theta"].copy(), info, "all", False)
    pop = ko.mc_points(1_300_003, 6, seed=5)
    r = akmcs_reductions(pop, info)
    o = kb.predict_arrays(pop, info, ("pred", "s"))
    G, sg = o["pred"], o["s"]
    U = np.abs(G) / sg
    assert r["minUloc"] == int(np.argmin(U)) and r["minU"] == float(np.min(U))
    n = len(pop)
    pf = np.count_nonzero(G <= 0) / n
    assert 0.0 < pf < 1.0 and r["Pf"] == pf and r["nGless"] == np.count_nonzero(G <= 0)
    assert r["cov"] == float(np.sqrt((1 - pf) / (pf * n)))
    pfp, pfn = np.count_nonzero(G - 1.96 * sg <= 0) / n, np.count_nonzero(G + 1.96 * sg <= 0) / n
    assert r["stop_criteria"] == (pfp - pfn) / pf
    # a population that never fails: the reference's sentinel values
    safe = akmcs_reductions(pop[:1000] * 0.0, ko_shift(info, +50.0, kb), )
    assert safe["Pf"] == 0 and safe["cov"] == 1000 and safe["stop_criteria"] == 100
    # the population resident in HBM across iterations (kb200_akmcs_reduce_dev): identical scalars, no upload per call
    from kadal_b200.reliability import DevicePopulation
    dp = DevicePopulation(pop)
    for _ in range(2):
        rd = akmcs_reductions(dp, info)
        assert rd == r
    assert np.array_equal(dp.rows(r["minUloc"]), pop[r["minUloc"]])
    import torch
    rt = akmcs_reductions(torch.from_numpy(pop[:200_001]).cuda(), info)
    assert rt == akmcs_reductions(pop[:200_001], info)


def ko_shift(info, dy, kb):
    """Same model with the response shifted by dy (refit at the same theta)."""
    K = ko.make_kriginfo(info["X"], info["y"] + dy)
    kb.likelihood(np.array(info["Theta"], dtype=float), K, "all", False)
    return K
