"""Round-2 golden vectors from the LIVE KADAL reference (build container only):

    PYTHONPATH=/root/repo python tests/golden/make_golden_r2.py

Writes tests/golden/r2_options.npz: the configurations VERDICT r1 found untested --
  * TrendOrder 1 and 2 likelihood (reference `Kriging` class builds idx / F with
    trendfunction.py:7-41, 72-111; likelihood_func.py:183-193 runs with F of p columns), concentrated and trainvar,
    one 'all' call + a small population each.  Prediction with p > 1 raises in the reference (prediction.py:250).
  * iso_gaussian (likelihood_func.py:60-61; one theta for every dimension), 'all' + predictions + a 1-D population.
  * PoF with limittype '<' and '<=', and 'lcb' with KrigInfo['sigmalcb'] set (prediction.py:268: (npred,1)-(1,npred)
    broadcast -> (npred, npred)).
Inputs are stored next to the outputs so the GPU box needs nothing but this file.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader, kriging_oracle as ko  # noqa: E402


def truth_F(ref, x, KI):
    import ctypes
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libtruth_quad.so"))
    lib.kb_truth_nll_quad_F.restype = ctypes.c_double
    lib.kb_truth_nll_quad_F.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p]
    A = copy.deepcopy(KI)
    ref.likelihood(x.copy(), A, "all", False)
    n = A["Psi"].shape[0]
    M = np.ascontiguousarray(A["Psi"] + np.eye(n) * 10.0 ** A["nugget"])
    F = np.ascontiguousarray(A["F"], dtype=float)
    y = np.ascontiguousarray(np.asarray(A["y"], dtype=float).ravel())
    return lib.kb_truth_nll_quad_F(M.ctypes.data, y.ctypes.data, F.ctypes.data, n, F.shape[1], None)


def trend_cases(ref, out):
    """TrendOrder 1 at d = 3 (p = 4 trend columns) and TrendOrder 2 at d = 2 (p = 6): the C ABI carries up to 7 trend
    columns next to y through the fused forward substitution (kb200_model_create)."""
    rng = np.random.default_rng(2025)
    n = 90
    lb3, ub3 = np.array([-2.0, 0.0, 1.0]), np.array([2.0, 1.0, 4.0])
    X3 = rng.uniform(lb3, ub3, (n, 3))
    y = (np.sin(X3[:, 0]) + 0.5 * X3[:, 1] ** 2 + 0.3 * X3[:, 2] + 0.2 * X3[:, 0] * X3[:, 2])[:, None] + 2.0
    for order, d in ((1, 3), (2, 2)):
        Xraw, lb, ub = X3[:, :d].copy(), lb3[:d], ub3[:d]
        x = np.array([0.2, -0.1, 0.3])[:d]
        pop = np.vstack([rng.uniform(-1, 1, (12, d)), rng.uniform(-3, 3, (8, d))])
        K = dict(X=Xraw, y=y, nvar=d, nsamp=n, nrestart=1, ub=ub, lb=lb, kernel=["gaussian"], TrendOrder=order,
                 nugget=-6, optimizer="lbfgsb")
        k = ref.Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
        KI = k.KrigInfo
        tag = "trend%d_" % order
        out[tag + "Xraw"], out[tag + "y"], out[tag + "lb"], out[tag + "ub"] = Xraw, y, lb, ub
        out[tag + "idx"] = np.asarray(KI["idx"], dtype=float)
        out[tag + "F"] = np.asarray(KI["F"], dtype=float)
        out[tag + "X_norm"] = np.asarray(KI["X_norm"], dtype=float)
        A = copy.deepcopy(KI)
        ref.likelihood(x.copy(), A, "all", False)
        out[tag + "x"] = x
        out[tag + "nll"] = A["NegLnLike"]
        out[tag + "nll_truth"] = truth_F(ref, x, KI)
        out[tag + "BE"] = np.asarray(A["BE"], dtype=float)
        out[tag + "s2"] = np.asarray(A["SigmaSqr"], dtype=float).reshape(-1)
        out[tag + "U"] = np.ascontiguousarray(A["U"])
        out[tag + "pop"] = pop
        out[tag + "pop_nll"] = np.array([ref.likelihood(p.copy(), copy.deepcopy(KI), "default", False) for p in pop])
        # binary128 value of the same formula on the same float64 Psi and F (oracle/truth/nll_quad.c): with p > 1 the
        # reference's own float64 result is up to 7e-9 away from it (TrendOrder 2, candidate 7)
        out[tag + "pop_truth"] = np.array([truth_F(ref, p, KI) for p in pop])
        # trainvar: x carries log10 sigma^2 behind theta (likelihood_func.py:74-79, 197-199)
        B = copy.deepcopy(KI)
        xt = np.r_[x, 0.4]
        ref.likelihood(xt.copy(), B, "all", True)
        out[tag + "tv_x"] = xt
        out[tag + "tv_nll"] = B["NegLnLike"]
        out[tag + "tv_BE"] = np.asarray(B["BE"], dtype=float)
        out[tag + "tv_s2"] = np.asarray(B["SigmaSqr"], dtype=float).reshape(-1)
        print(tag, "p =", KI["F"].shape[1], "nll", A["NegLnLike"], "trainvar nll", B["NegLnLike"])


def iso_case(ref, out):
    rng = np.random.default_rng(31)
    n, d = 70, 4
    X = rng.uniform(-1, 1, (n, d))
    y = np.cos(X.sum(1, keepdims=True)) + 0.2 * X[:, 1:2] + 3.0
    info = ko.make_kriginfo(X, y, kernel=("iso_gaussian",))
    xp = rng.uniform(-1, 1, (40, d))
    A = copy.deepcopy(info)
    ref.likelihood(0.15, A, "all", False)                 # scalar x, as minimize_scalar passes it (kriging_model.py:210-213)
    p = ref.prediction(xp, A, ["pred", "SSqr", "s"])
    grid = np.linspace(-1.5, 1.5, 16)
    out.update(iso_X=X, iso_y=y, iso_xp=xp, iso_x=0.15, iso_nll=A["NegLnLike"], iso_BE=np.asarray(A["BE"]),
               iso_s2=np.asarray(A["SigmaSqr"]).reshape(-1), iso_Theta=np.asarray(A["Theta"], dtype=float),
               iso_pred=np.asarray(p[0]).reshape(-1), iso_SSqr=np.asarray(p[1]).reshape(-1), iso_s=np.asarray(p[2]).reshape(-1),
               iso_grid=grid,
               iso_grid_nll=np.array([ref.likelihood(float(g), copy.deepcopy(info), "default", False) for g in grid]))
    print("iso nll", A["NegLnLike"], "Theta", A["Theta"])


def pof_lcb_case(ref, out):
    rng = np.random.default_rng(57)
    n, d = 60, 3
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X.sum(1, keepdims=True)) + X[:, :1] ** 2 + 1.0
    xp = rng.uniform(-1, 1, (25, d))
    x = np.array([-0.5, -0.6, -0.4])            # short length scales: s comparable to |f - limit|, erf not saturated
    info = ko.make_kriginfo(X, y, kernel=("matern52",))
    ref.likelihood(x.copy(), info, "all", False)
    out.update(pl_X=X, pl_y=y, pl_xp=xp, pl_x=x, pl_limit=1.4, pl_sigmalcb=2.5)
    for lt, tag in (("<", "lt"), ("<=", "le"), (">", "gt")):
        A = copy.deepcopy(info)
        A["limit"] = 1.4
        A["limittype"] = lt
        out["pl_pof_" + tag] = np.asarray(ref.prediction(xp, A, "PoF")).reshape(-1)
    A = copy.deepcopy(info)
    A["sigmalcb"] = 2.5
    lcb = ref.prediction(xp, A, "LCB")
    out["pl_lcb"] = np.asarray(lcb)
    ebe = ref.prediction(xp, A, "ebe")
    out["pl_ebe"] = np.asarray(ebe)
    print("lcb shape", np.shape(lcb), "ebe shape", np.shape(ebe), "pof<", out["pl_pof_lt"][:3])


if __name__ == "__main__":
    ref = ref_loader.load()
    out = {}
    trend_cases(ref, out)
    iso_case(ref, out)
    pof_lcb_case(ref, out)
    np.savez_compressed(os.path.join(HERE, "r2_options.npz"), **out)
    print("r2_options.npz written:", len(out), "arrays")
