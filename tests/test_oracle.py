"""CPU: the oracle restatement against the golden vectors of the live reference, against
SURVEY.md Appendix C's 17-digit values, and (when /root/reference is present) against the
reference itself, bit for bit."""
import copy

import numpy as np
import pytest

from conftest import relerr
import cases
from oracle import kriging_oracle as ko, ref_loader

# SURVEY.md Appendix C (values printed by the unmodified reference)
APPENDIX_C_NLL = {
    "gaussian": -0.5330237149104136, "exponential": -24.556348761701216, "matern32": -25.566083849230235,
    "matern52": -22.37917465705773, "trainvar": 250.18106627066476, "tuned": -2.9510607602839656,
    "composite": -22.751351305346457, "kpls": 8.522583301819068,
}


@pytest.mark.parametrize("name", cases.APPENDIX_C_CASES)
def test_oracle_matches_golden_appendix_c(golden_c, name):
    g = golden_c
    info, x, tv = cases.appendix_c_info(g, name)
    ko.likelihood(x, info, "all", tv)
    assert abs(float(g[name + "_nll"]) - APPENDIX_C_NLL[name]) < 1e-9 * max(1, abs(APPENDIX_C_NLL[name]))
    assert relerr(info["NegLnLike"], g[name + "_nll"]) < 1e-12
    assert relerr(info["BE"], g[name + "_BE"]) < 1e-11
    assert relerr(np.asarray(info["SigmaSqr"]).ravel(), g[name + "_s2"], 1e-300) < 1e-11
    assert relerr(info["U"], g[name + "_U"]) < 1e-12
    assert relerr(info["Psi"], g[name + "_Psi"]) < 1e-14
    out = ko.prediction(g["xp"], info, cases.PT)
    for k, v in zip(cases.PT, out):
        assert relerr(np.asarray(v).ravel(), g[name + "_" + k], 1e-300) < 1e-9, k


def test_oracle_matches_golden_c3(golden_c3):
    g = golden_c3
    info, xp = cases.c3_info(g)
    nll = ko.likelihood_population(g["pop"], copy.deepcopy(info), False, check_pd=False)
    assert relerr(nll, g["pop_nll"]) < 1e-12
    assert int(np.argmin(nll)) == int(np.argmin(g["pop_nll"]))
    ko.likelihood(g["fit_x"].copy(), info, "all", False, check_pd=False)
    out = ko.prediction(xp, info, cases.PT + ["PoF"])
    for k, v in zip(cases.PT + ["PoF"], out):
        assert relerr(np.asarray(v).ravel(), g[k], 1e-300) < 1e-9, k


def test_oracle_matches_golden_std(golden_std):
    g = golden_std
    info = cases.std_info(g)
    ko.likelihood(g["x"].copy(), info, "all", False)
    assert relerr(info["NegLnLike"], g["nll"]) < 1e-12
    out = ko.prediction(g["xp"], info, ["pred", "SSqr", "s"])
    for k, v in zip(["pred", "SSqr", "s"], out):
        assert relerr(np.asarray(v).ravel(), g[k], 1e-300) < 1e-9, k


def test_oracle_c2_known_answer(golden_c2):
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    v = ko.likelihood(w["kat_x"], info, "default", False, check_pd=False)
    assert abs(float(golden_c2["kat_nll"]) - (-345.9939996418981)) < 1e-10
    assert abs(v - float(golden_c2["kat_nll"])) < 1e-9 * 346


def test_oracle_c2_full_population_golden(golden_c2_full, golden_c2):
    """c2_full.npz holds the live reference on all 2 x 512 candidates; spot-check the oracle on a few of each
    population (a full pass is 20 CPU-minutes) and the file against the older anchors."""
    g = golden_c2_full
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    assert relerr(g["nll_popB"][golden_c2["idxB"]], golden_c2["nllB"]) < 1e-13
    assert relerr(g["nll_popA"][golden_c2["idxA"]], golden_c2["nllA"]) < 1e-13
    for pop, idx in (("popA", [100, 316]), ("popB", [114, 178, 476])):
        v = ko.likelihood_population(w[pop][idx], copy.deepcopy(info), False, check_pd=False)
        assert relerr(v, g["nll_" + pop][idx]) < 1e-12
    assert int(np.argmin(g["nll_popB"])) == 178 and g["cond_popB"].max() > 9e8


def test_c2_truth_fixture(golden_c2_truth, golden_c2_full):
    """The reference's float64 result against the long-double value of its own formula: the reference itself
    is 0.01 .. 1.4 cond ulp away from exact -- the scale of any parity bound at cond ~ 1e9."""
    t = golden_c2_truth
    assert np.array_equal(t["ref"], golden_c2_full["nll_popB"][t["idx"]]) or relerr(t["ref"], golden_c2_full["nll_popB"][t["idx"]]) < 1e-13
    e = np.abs(t["ref"] - t["truth"]) / np.maximum(1, np.abs(t["truth"])) / (t["cond"] * 2.0 ** -53)
    assert 114 in t["idx"] and e.max() < 2.0 and e.max() > 0.05


def test_oracle_errors():
    info = ko.make_kriginfo(np.zeros((4, 2)), np.zeros((4, 1)))
    with pytest.raises(np.linalg.LinAlgError):  # duplicate rows + 0 nugget-free? keep nugget tiny
        info["nugget"] = -300
        ko.likelihood(np.zeros(2), info, "default", False, check_pd=False)
    with pytest.raises(ValueError):
        ko.calckernel(np.zeros((2, 1)), np.zeros((2, 1)), np.ones(1), 1, type="cubic")
    with pytest.raises(TypeError):
        ko.likelihood(np.zeros(2), ko.make_kriginfo(np.eye(2), np.ones((2, 1))), "bogus", False)


@pytest.mark.skipif(not ref_loader.available(), reason="live reference only exists in the build container")
@pytest.mark.parametrize("name", cases.APPENDIX_C_CASES)
def test_oracle_bit_identical_to_live_reference(golden_c, name):
    ref = ref_loader.load()
    info, x, tv = cases.appendix_c_info(golden_c, name)
    A, B = copy.deepcopy(info), copy.deepcopy(info)
    ref.likelihood(x.copy(), A, "all", tv)
    ko.likelihood(x.copy(), B, "all", tv)
    assert np.array_equal(A["Psi"], B["Psi"]) and np.array_equal(A["U"], B["U"])
    assert A["NegLnLike"] == B["NegLnLike"]
    pa = ref.prediction(golden_c["xp"], A, cases.PT)
    pb = ko.prediction(golden_c["xp"], B, cases.PT)
    for a, b in zip(pa, pb):
        assert np.array_equal(np.asarray(a), np.asarray(b))
