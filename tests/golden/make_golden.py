"""Generate tests/golden/*.npz from the LIVE KADAL reference (build container only).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

The reference ships no tests or golden vectors for the Kriging path, so the
vectors pinned here are outputs of the unmodified reference functions
(kadal/surrogate_models/supports/likelihood_func.py:likelihood, prediction.py:prediction)
imported from /root/reference through oracle/ref_loader.py.  Inputs are regenerated from the
seeds stored in each file, outputs are stored with full precision.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader, kriging_oracle as ko  # noqa: E402

PT = ["pred", "SSqr", "s", "EI", "PoI"]


def appendix_c(ref):
    """n=64, d=4, all kernels / modes."""
    rng = np.random.default_rng(12345)
    n, d = 64, 4
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(2 * X.sum(1, keepdims=True)) + X[:, :1] ** 2
    xp = rng.uniform(-1, 1, (3, d))
    pls = rng.standard_normal((d, 2))
    x = np.array([0.1, -0.2, 0.3, 0.0])
    cases = {
        "gaussian": (dict(kernel=("gaussian",)), x, False),
        "exponential": (dict(kernel=("exponential",)), x, False),
        "matern32": (dict(kernel=("matern32",)), x, False),
        "matern52": (dict(kernel=("matern52",)), x, False),
        "trainvar": (dict(kernel=("gaussian",)), np.r_[x, 0.5], True),
        "tuned": (dict(kernel=("gaussian",)), np.r_[x, -4.0], False),
        "composite": (dict(kernel=("gaussian", "matern32", "exponential")), np.r_[x, .5, .3, .2], False),
        "kpls": (dict(kernel=("gaussian",), pls=pls, n_princomp=2), np.array([0.1, -0.2]), False),
    }
    out = dict(X=X, y=y, xp=xp, pls=pls, names=np.array(list(cases)))
    for name, (kw, xx, tv) in cases.items():
        A = ko.make_kriginfo(X, y, **kw)
        ref.likelihood(xx.copy(), A, "all", tv)
        p = ref.prediction(xp, A, PT)
        out[name + "_x"] = xx
        out[name + "_trainvar"] = tv
        out[name + "_kernel"] = np.array(kw["kernel"])
        out[name + "_nll"] = A["NegLnLike"]
        out[name + "_BE"] = A["BE"]
        out[name + "_s2"] = np.asarray(A["SigmaSqr"]).reshape(-1)
        out[name + "_U"] = np.ascontiguousarray(A["U"])
        out[name + "_Psi"] = A["Psi"]
        for k, v in zip(PT, p):
            out[name + "_" + k] = np.asarray(v).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "appendix_c.npz"), **out)
    print("appendix_c.npz", {k: float(out[k + "_nll"]) for k in cases})


def c3_like(ref):
    """n=200, d=6 (BASELINE config 3 shape): a 48-candidate population + bulk predictions,
    'default' normalisation with non-trivial lb/ub, PoF with a limit."""
    w = ko.workload("C3")
    X, y = w["X"], w["y"]
    n, d = X.shape
    lb, ub = np.full(d, -2.0), np.linspace(1.5, 3.0, d)
    Xraw = (X / 2 + 0.5) * (ub - lb) + lb
    info = ko.make_kriginfo(X, y, lb=lb, ub=ub)
    info["X"] = Xraw
    info["limit"] = 3.2
    info["limittype"] = ">="
    rng = np.random.default_rng(42)
    pop = np.vstack([rng.uniform(-5, 5, (16, d)), rng.uniform(-1, 1, (16, d)), rng.uniform(0.0, 1.0, (16, d))])
    nll = np.array([ref.likelihood(p.copy(), copy.deepcopy(info), "default", False) for p in pop])
    fit = copy.deepcopy(info)
    ref.likelihood(np.full(d, 0.2), fit, "all", False)
    xp = (ko.mc_points(2000, d, seed=5) / 2 + 0.5) * (ub - lb) + lb
    p = ref.prediction(xp, fit, PT + ["PoF"])
    out = dict(lb=lb, ub=ub, pop=pop, pop_nll=nll, fit_x=np.full(d, 0.2), fit_nll=fit["NegLnLike"],
               fit_BE=fit["BE"], fit_s2=np.asarray(fit["SigmaSqr"]).reshape(-1), xp_seed=5, npred=2000,
               cond=np.linalg.cond(fit["Psi"] + 1e-6 * np.eye(n)))
    for k, v in zip(PT + ["PoF"], p):
        out[k] = np.asarray(v).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "c3_like.npz"), **out)
    print("c3_like.npz nll range", nll.min(), nll.max(), "cond", out["cond"])


def std_norm(ref):
    """standtype='std' through the reference's own Kriging class (n=60, d=3, matern52)."""
    rng = np.random.default_rng(77)
    n, d = 60, 3
    lb, ub = np.array([-5.0, 0.0, 10.0]), np.array([5.0, 2.0, 30.0])
    Xraw = rng.uniform(lb, ub, (n, d))
    y = (np.sin(Xraw[:, 0]) + Xraw[:, 1] ** 2 + 0.1 * Xraw[:, 2])[:, None]
    K = dict(X=Xraw, y=y, nvar=d, nsamp=n, nrestart=1, ub=ub, lb=lb, kernel=["matern52"], TrendOrder=0,
             nugget=-6, optimizer="lbfgsb")
    k = ref.Kriging(K, standardization=True, standtype="std", normy=False, trainvar=False)
    x = np.array([0.3, -0.1, 0.4])
    k.KrigInfo = ref.likelihood(x, k.KrigInfo, "all", False)
    xp = rng.uniform(lb, ub, (500, d))
    p = ref.prediction(xp, k.KrigInfo, ["pred", "SSqr", "s"])
    KI = k.KrigInfo
    np.savez_compressed(os.path.join(HERE, "std_norm.npz"), Xraw=Xraw, y=y, lb=lb, ub=ub, x=x, xp=xp,
                        X_norm=KI["X_norm"], X_mean=KI["X_mean"], X_std=KI["X_std"], F=KI["F"],
                        nll=KI["NegLnLike"], BE=KI["BE"], s2=np.asarray(KI["SigmaSqr"]).reshape(-1),
                        pred=np.asarray(p[0]).reshape(-1), SSqr=np.asarray(p[1]).reshape(-1),
                        s=np.asarray(p[2]).reshape(-1))
    print("std_norm.npz nll", KI["NegLnLike"])


def c2_anchor(ref):
    """n=1000, d=10 (BASELINE config 2): known-answer + first candidates of both populations."""
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    kat = ref.likelihood(w["kat_x"].copy(), copy.deepcopy(info), "default", False)
    idxA, idxB = np.arange(6), np.arange(10)
    nA = np.array([ref.likelihood(w["popA"][i].copy(), copy.deepcopy(info), "default", False) for i in idxA])
    nB = np.array([ref.likelihood(w["popB"][i].copy(), copy.deepcopy(info), "default", False) for i in idxB])
    np.savez_compressed(os.path.join(HERE, "c2_anchor.npz"), kat_nll=kat, idxA=idxA, nllA=nA, idxB=idxB, nllB=nB)
    print("c2_anchor.npz KAT", repr(kat))


def _c2_full_worker(args):
    name, lo, hi = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "2"
    ref = ref_loader.load()
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    n = w["X"].shape[0]
    nll, cond = np.empty(hi - lo), np.empty(hi - lo)
    for k, i in enumerate(range(lo, hi)):
        A = copy.deepcopy(info)
        ref.likelihood(w[name][i].copy(), A, "all", False)
        nll[k] = A["NegLnLike"]
        ev = np.linalg.eigvalsh(A["Psi"] + np.eye(n) * 1e-6)
        cond[k] = ev[-1] / ev[0]
    return name, lo, nll, cond


def c2_full():
    """n=1000, d=10 (BASELINE config 2): the live reference on EVERY candidate of both 512-candidate
    populations, with cond(Psi + eps I) for the tolerance model of SURVEY.md Appendix E."""
    from multiprocessing import get_context
    jobs = [(name, lo, lo + 32) for name in ("popA", "popB") for lo in range(0, 512, 32)]
    out = dict(nll_popA=np.empty(512), nll_popB=np.empty(512), cond_popA=np.empty(512), cond_popB=np.empty(512))
    with get_context("spawn").Pool(4) as pool:
        for name, lo, nll, cond in pool.imap_unordered(_c2_full_worker, jobs):
            out["nll_" + name][lo:lo + 32] = nll
            out["cond_" + name][lo:lo + 32] = cond
            print(name, lo, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "c2_full.npz"), **out)
    print("c2_full.npz", out["nll_popB"].min(), out["cond_popB"].max(), out["cond_popA"].max())


def ehvi_fronts():
    """Deterministic Pareto fronts + candidates of the EHVI goldens (shared with the tests)."""
    rng = np.random.default_rng(2024)
    cases = []
    for k in (1, 2, 5, 13, 40):
        a = np.sort(rng.uniform(0, 1, k))
        b = np.sort(rng.uniform(0, 1, k))[::-1]
        P = np.c_[a, b][rng.permutation(k)]          # non-dominated, rows shuffled
        r = np.array([1.2, 1.3])
        mu = np.vstack([rng.uniform(-0.2, 1.4, (10, 2)), P[:2] + 1e-3, [[1.5, 1.6]]])   # inside, on the front, beyond r
        s = np.vstack([rng.uniform(1e-3, 0.5, (10, 2)), np.full((len(mu) - 10, 2), 0.05)])
        cases.append((P, r, mu, s))
    return cases


def ehvi2d():
    """exi2d outputs of the live reference (kadal/optim_tools/ehvi/exi2d.py) — SURVEY.md §8 f1."""
    import warnings
    ref_loader._ensure_path()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from kadal.optim_tools.ehvi.exi2d import exi2d
    out = {}
    for c, (P, r, mu, s) in enumerate(ehvi_fronts()):
        out["P%d" % c], out["r%d" % c], out["mu%d" % c], out["s%d" % c] = P, r, mu, s
        out["v%d" % c] = np.array([exi2d(P, r, mu[i], s[i]) for i in range(len(mu))])
    out["ncases"] = len(ehvi_fronts())
    np.savez_compressed(os.path.join(HERE, "ehvi2d.npz"), **out)
    print("ehvi2d.npz", [float(out["v%d" % c].max()) for c in range(out["ncases"])])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ehvi":
        ehvi2d()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2full":
        c2_full()
        sys.exit(0)
    ref = ref_loader.load()
    appendix_c(ref)
    c3_like(ref)
    std_norm(ref)
    c2_anchor(ref)
    ehvi2d()
