"""Shared builders for the parity cases (inputs regenerated from seeds, cf. make_golden.py)."""
import numpy as np

from oracle import kriging_oracle as ko


def appendix_c_info(g, name):
    kw = dict(kernel=tuple(str(k) for k in g[name + "_kernel"]))
    if name == "kpls":
        kw.update(pls=g["pls"], n_princomp=2)
    return ko.make_kriginfo(g["X"], g["y"], **kw), g[name + "_x"].copy(), bool(g[name + "_trainvar"])


def c3_info(g):
    w = ko.workload("C3")
    X, y = w["X"], w["y"]
    lb, ub = g["lb"], g["ub"]
    info = ko.make_kriginfo(X, y, lb=lb, ub=ub)
    info["X"] = (X / 2 + 0.5) * (ub - lb) + lb
    info["limit"] = 3.2
    info["limittype"] = ">="
    xp = (ko.mc_points(int(g["npred"]), X.shape[1], seed=int(g["xp_seed"])) / 2 + 0.5) * (ub - lb) + lb
    return info, xp


def std_info(g):
    n, d = g["Xraw"].shape
    return dict(nvar=d, nsamp=n, X=g["Xraw"], X_norm=g["X_norm"], y=g["y"], F=g["F"], kernel=["matern52"],
                nkernel=1, standardization=True, normtype="std", norm_y=False, type="kriging",
                n_princomp=False, nugget=-6, idx=np.zeros((1, d)), TrendOrder=0, lb=g["lb"], ub=g["ub"],
                X_mean=g["X_mean"], X_std=g["X_std"])


def trend_info(g, order):
    """KrigInfo of the TrendOrder 1 / 2 goldens (r2_options.npz): X_norm, idx and F exactly as the reference's
    Kriging.standardize wrote them (kriging_model.py:99-164)."""
    t = "trend%d_" % order
    n, d = g[t + "Xraw"].shape
    return dict(nvar=d, nsamp=n, X=g[t + "Xraw"], X_norm=g[t + "X_norm"], y=g[t + "y"], F=g[t + "F"], idx=g[t + "idx"],
                kernel=["gaussian"], nkernel=1, standardization=True, normtype="default", norm_y=False, type="kriging",
                n_princomp=False, nugget=-6, TrendOrder=order, lb=g[t + "lb"], ub=g[t + "ub"])


def iso_info(g):
    return ko.make_kriginfo(g["iso_X"], g["iso_y"], kernel=("iso_gaussian",))


def pof_lcb_info(g):
    return ko.make_kriginfo(g["pl_X"], g["pl_y"], kernel=("matern52",))


APPENDIX_C_CASES = ["gaussian", "exponential", "matern32", "matern52", "trainvar", "tuned", "composite", "kpls"]
PT = ["pred", "SSqr", "s", "EI", "PoI"]
