"""Drop-in for ``kadal/surrogate_models/kriging_model.py``: the ``Kriging`` class with the
reference's constructor, ``standardize`` / ``train`` / ``predict`` / ``parallelopt`` methods,
``tune_hyperparameters`` and ``kriginfocheck`` (same names, argument meaning, KrigInfo keys
and error behaviour — kriging_model.py:51-97, 99-164, 166-258, 299-325, 327-392, 395-458,
461-574).

What changes is how the hyper-parameter search evaluates the likelihood.  The reference
hands ``likelihood`` to scipy / cma, which call it one candidate at a time.  Here every
optimiser asks for *batches*:

* ``lbfgsb`` / ``slsqp`` — value and forward-difference gradient of one iterate are one batch
  of ``nbhyp + 1`` candidates (the reference gets the same stencil from scipy's internal
  finite differences, ``options={'eps': 1e-3}``, kriging_model.py:433-436);
* ``cmaes`` — a whole CMA-ES generation is one batch (:mod:`kadal_b200.cmaes`);
* all ``nrestart`` restarts advance in lock step and their batches are concatenated, so one
  ``kb200_lik_batch`` call serves every restart (the reference offers a process pool over
  restarts instead, :366-371 — ``n_cpu`` / ``pool`` are accepted and ignored because CUDA
  contexts do not survive ``fork``).

All likelihood arithmetic runs on the GPU; this module is control logic only.
"""
import threading

import numpy as np
from scipy.optimize import fmin_cobyla, minimize

from . import cmaes
from .likelihood_func import likelihood, likelihood_batch
from .prediction import prediction
from .sampling import sampling, standardize
from .trendfunction import compute_regression_mat, polytruncation

PENALTY = 10000.0   # the reference's NegLnLike for a failed evaluation (likelihood_func.py:207)


# ------------------------------------------------------------------ batched objective
class BatchedObjective:
    """``f(X[k, nbhyp]) -> nll[k]`` on the GPU, counting evaluations and batches.
    Non positive definite candidates get the reference's failure penalty (10000) so that
    a finite-difference stencil or a CMA-ES generation survives them."""

    def __init__(self, KrigInfo, trainvar):
        self.KrigInfo = KrigInfo
        self.trainvar = trainvar
        self.nevals = 0
        self.nbatches = 0

    def __call__(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        nll, info = likelihood_batch(X, self.KrigInfo, self.trainvar)
        self.nevals += X.shape[0]
        self.nbatches += 1
        bad = (info != 0) | ~np.isfinite(nll)
        if np.any(bad):
            nll = np.where(bad, PENALTY, nll)
        return nll


class _Lockstep:
    """Rendezvous of several optimiser threads: each submits the batch it needs next; when
    every live worker has submitted, the batches are concatenated into ONE GPU call."""

    def __init__(self, batch_fun, nworkers):
        self.fun = batch_fun
        self.cv = threading.Condition()
        self.live = nworkers
        self.pending = {}
        self.results = {}
        self.error = None

    def _flush(self):
        ids = list(self.pending)
        blocks = [self.pending[i] for i in ids]
        try:
            f = np.asarray(self.fun(np.vstack(blocks)), dtype=float).reshape(-1)
            k = 0
            for i, b in zip(ids, blocks):
                self.results[i] = f[k:k + len(b)]
                k += len(b)
        except BaseException as e:   # propagate to every waiting worker
            self.error = e
            for i in ids:
                self.results[i] = None
        self.pending.clear()
        self.cv.notify_all()

    def evaluate(self, wid, X):
        with self.cv:
            self.pending[wid] = np.atleast_2d(X)
            if len(self.pending) >= self.live:
                self._flush()
            while wid not in self.results:
                self.cv.wait()
            out = self.results.pop(wid)
            if out is None:
                raise self.error
            return out

    def done(self, wid):
        with self.cv:
            self.live -= 1
            if self.pending and len(self.pending) >= self.live:
                self._flush()


def _fd_steps(x, lb, ub, eps):
    """Forward-difference steps with scipy's bound handling for the '2-point' scheme
    (scipy.optimize._numdiff._adjust_scheme_to_bounds, '1-sided'): flip a step that would
    leave the box when the other side has room, otherwise shrink it to the larger side."""
    h = np.full(x.shape, float(eps))
    if lb is None:
        return h
    lower, upper = x - lb, ub - x
    xh = x + h
    violated = (xh < lb) | (xh > ub)
    fitting = np.abs(h) <= np.maximum(lower, upper)
    h[violated & fitting] *= -1
    fwd = (upper >= lower) & ~fitting
    h[fwd] = upper[fwd]
    bwd = (upper < lower) & ~fitting
    h[bwd] = -lower[bwd]
    return h


def _value_and_grad(evaluate, x, lb, ub, eps):
    """One batch: [x; x + h_i e_i] -> f(x), forward-difference gradient."""
    x = np.asarray(x, dtype=float)
    h = _fd_steps(x, lb, ub, eps)
    X = np.vstack([x, x + np.diag(h)])
    f = evaluate(X)
    dx = X[1:, :].diagonal() - x
    g = np.where(dx != 0, (f[1:] - f[0]) / np.where(dx != 0, dx, 1.0), 0.0)
    return float(f[0]), g


def _run_scipy(method, evaluate, x0, optimbound):
    lb = ub = None
    if optimbound is not None and method != "cobyla":
        ob = np.asarray(optimbound, dtype=float)
        lb, ub = ob[:, 0], ob[:, 1]
    if method == "lbfgsb":
        # kriging_model.py:433-436: L-BFGS-B, absolute FD step 1e-3, bounds
        res = minimize(lambda x: _value_and_grad(evaluate, x, lb, ub, 1e-3), x0, jac=True,
                       method="L-BFGS-B", bounds=optimbound)
        return res.x, float(res.fun)
    if method == "slsqp":
        # kriging_model.py:441-444: SLSQP with scipy's default FD step sqrt(eps_mach)
        step = float(np.sqrt(np.finfo(float).eps))
        res = minimize(lambda x: _value_and_grad(evaluate, x, lb, ub, step), x0, jac=True,
                       method="SLSQP", bounds=optimbound)
        return res.x, float(res.fun)
    if method == "cobyla":
        # kriging_model.py:449-452: derivative free, inherently one candidate per step
        res = fmin_cobyla(lambda x, *a: float(evaluate(np.atleast_2d(x))[0]), x0, optimbound,
                          rhobeg=0.5, rhoend=1e-4, args=(None, None, None))
        return np.asarray(res), float(evaluate(np.atleast_2d(res))[0])
    raise KeyError("%s in KrigInfo['Optimizer'] is not recognised." % method)


def tune_hyperparameters(KrigInfo, xhyp_ii, trainvar, ubhyp=None, lbhyp=None,
                         sigmacmaes=None, scaling=None, optimbound=None):
    """Best hyper-parameters from one starting point (kriging_model.py:395-458)."""
    xs = np.atleast_2d(np.asarray(xhyp_ii, dtype=float))
    bx, bf, _ = _tune_many(KrigInfo, xs, trainvar, ubhyp, lbhyp, sigmacmaes, scaling, optimbound)
    return bx[0], bf[0]


def _tune_many(KrigInfo, xhyp, trainvar, ubhyp, lbhyp, sigmacmaes, scaling, optimbound, seed=None):
    """All restarts in lock step; returns (bestx[nrestart, nbhyp], nll[nrestart], objective)."""
    opt = KrigInfo["optimizer"]
    obj = BatchedObjective(KrigInfo, trainvar)
    xhyp = np.atleast_2d(np.asarray(xhyp, dtype=float))
    nres = xhyp.shape[0]
    if opt == "cmaes":
        for name, p in (("ubhyp", ubhyp), ("lbhyp", lbhyp), ("sigmacmaes", sigmacmaes), ("scaling", scaling)):
            if p is None:
                raise ValueError("%s must be set if optimizer is cmaes." % name)
        bx, bf, _ = cmaes.fmin_lockstep(obj, xhyp, sigmacmaes, bounds=(np.asarray(lbhyp), np.asarray(ubhyp)),
                                        scaling=scaling, seed=seed)
        return bx, bf, obj
    if opt not in ("lbfgsb", "slsqp", "cobyla"):
        # 'ga' passes kriginfocheck in the reference as well (kriging_model.py:490) but is only wired up for
        # nbhyp == 1 (:219-221); with more hyper-parameters its tune_hyperparameters ends in this same KeyError (:455-458)
        raise KeyError("%s in KrigInfo['Optimizer'] is not recognised." % opt)
    if optimbound is None:
        raise ValueError("optimbound must be set if optimizer is %s." % opt)
    if nres == 1:
        x, f = _run_scipy(opt, obj, xhyp[0], optimbound)
        return np.atleast_2d(x), np.array([f]), obj
    gate = _Lockstep(obj, nres)
    bestx = np.zeros_like(xhyp)
    bestf = np.full(nres, np.inf)
    errors = []

    def worker(i):
        try:
            x, f = _run_scipy(opt, lambda X: gate.evaluate(i, X), xhyp[i], optimbound)
            bestx[i], bestf[i] = x, f
        except BaseException as e:
            errors.append(e)
        finally:
            gate.done(i)

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(nres)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return bestx, bestf, obj


def _grid_zoom(obj, lo, hi, npts=65, rounds=5):
    """1-D bounded minimisation by repeated batched grids (nbhyp == 1).  The reference uses
    ``minimize_scalar(method='golden', bounds=...)`` (kriging_model.py:210-213), which modern
    scipy rejects; a grid of 65 candidates per batch, zoomed 5 times
    around the best, resolves the optimum to (hi-lo)/32**5 with 5 GPU calls."""
    best_x, best_f = None, np.inf
    for _ in range(rounds):
        xs = np.linspace(lo, hi, npts)
        f = obj(xs[:, None])
        i = int(np.argmin(f))
        if f[i] < best_f:
            best_x, best_f = xs[i], float(f[i])
        step = xs[1] - xs[0]
        lo, hi = max(lo, xs[i] - step), min(hi, xs[i] + step)
    return best_x, best_f


# ------------------------------------------------------------------ the model class
class Kriging:
    """Kriging model with the reference's interface (kriging_model.py:16-97).

    KrigInfo (dict) carries the model; required keys: 'X', 'y', 'lb', 'ub'; optional:
    'nugget' (log10; a [lo, hi] list makes it tuned), 'TrendOrder', 'kernel' (str or list:
    gaussian | exponential | matern32 | matern52 | iso_gaussian), 'nrestart', 'optimizer'
    (lbfgsb | cmaes | cobyla | slsqp | ga), 'limit' / 'limittype'."""

    def __init__(self, KrigInfo, ub=5, lb=-5, standardization=False, standtype="default", normy=False,
                 trainvar=False, inherit=False):
        KrigInfo["nvar"] = np.size(KrigInfo["X"], 1)
        KrigInfo["nsamp"] = np.size(KrigInfo["X"], 0)
        self.nbhyp = KrigInfo["nvar"] + 1 if trainvar is True else KrigInfo["nvar"]
        self.trainvar = trainvar
        self.normy = normy
        self.standardization = standardization
        self.standtype = standtype
        self.Y = KrigInfo["y"]
        self.X = KrigInfo["X"]
        self.lb = lb
        self.ub = ub
        self.train_stats = None
        if not inherit:
            self.type = "kriging"
            KrigInfo["type"] = self.type
            KrigInfo, scaling = kriginfocheck(KrigInfo, lb, ub, self.nbhyp)
            self.KrigInfo = KrigInfo
            self.scaling = scaling
            self.sigmacmaes = (ub - lb) / 5
            if self.standardization is True:
                self.standardize()

    def standardize(self):
        """Standardise the samples and build the regression matrix F (kriging_model.py:99-164)."""
        K = self.KrigInfo
        K["idx"] = polytruncation(K["TrendOrder"], K["nvar"], 1)
        nvar = K["nvar"]
        if self.standardization is True:
            if self.standtype.lower() == "default":
                K["normtype"] = "default"
                bound = np.vstack((-np.ones((1, nvar)), np.ones((1, nvar))))
                if self.normy is True:
                    rng = np.vstack((np.hstack((K["lb"], np.min(self.Y))), np.hstack((K["ub"], np.max(self.Y)))))
                    K["X_norm"], K["y_norm"] = standardize(K["X"], K["y"], type="default", norm_y=True, range=rng)
                    K["norm_y"] = True
                else:
                    K["X_norm"] = standardize(K["X"], K["y"], type="default", range=np.vstack((K["lb"], K["ub"])))
                    K["norm_y"] = False
            else:
                K["normtype"] = "std"
                if self.normy is True:
                    (K["X_norm"], K["y_norm"], K["X_mean"], K["y_mean"], K["X_std"], K["y_std"]) = \
                        standardize(K["X"], K["y"], type="std", norm_y=True)
                    K["norm_y"] = True
                else:
                    K["X_norm"], K["X_mean"], K["X_std"] = standardize(K["X"], K["y"], type="std")
                    K["norm_y"] = False
                bound = np.vstack((np.min(K["X_norm"], axis=0), np.max(K["X_norm"], axis=0)))
            K["standardization"] = True
            K["F"] = compute_regression_mat(K["idx"], K["X_norm"], bound, np.ones(nvar))
        else:
            K["standardization"] = False
            K["norm_y"] = False
            bound = np.vstack((np.min(K["X"], axis=0), np.max(K["X"], axis=0)))
            K["F"] = compute_regression_mat(K["idx"], K["X"], bound, np.ones(nvar))

    def _ensure_ready(self):
        # the reference leaves F / idx undefined when standardization=False (KeyError 'F',
        # SURVEY.md Appendix B); build them here so the raw-input mode is usable
        if "F" not in self.KrigInfo:
            self.standardize()

    def train(self, n_cpu=1, disp=True, pre_theta=None, pool=None, seed=None):
        """Multi-start hyper-parameter search, then ``likelihood(best, mode='all')``
        (kriging_model.py:166-258).  ``n_cpu`` / ``pool`` are accepted for compatibility;
        restarts run in lock step on the GPU instead of in a process pool."""
        self._ensure_ready()
        K = self.KrigInfo
        if disp:
            print("Begin train hyperparam.")
        if K["kernel"] == ["iso_gaussian"]:
            self.nbhyp = 1
        elif len(K["kernel"]) != 1 and "iso_gaussian" in K["kernel"]:
            raise NotImplementedError("Isotropic Gaussian kernel is not available for composite kernel")
        elif len(K["ubhyp"]) != self.nbhyp:
            self.nbhyp = len(K["ubhyp"])

        if K["nrestart"] < 1:
            xhyp = np.zeros((1, self.nbhyp))
        else:
            kind = "sobol" if self.nbhyp <= 40 else "sobolnew"
            _, xhyp = sampling(kind, len(K["ubhyp"]), K["nrestart"], result="real",
                               upbound=K["ubhyp"], lobound=K["lbhyp"])
        if pre_theta is not None:
            pre_theta = np.asarray(pre_theta, dtype=float)
            xhyp = np.random.rand(K["nrestart"] - 1, len(K["ubhyp"])) + pre_theta
            xhyp = np.vstack((pre_theta, xhyp))

        if self.nbhyp == 1:
            obj = BatchedObjective(K, self.trainvar)
            x1, neglnlikecand = _grid_zoom(obj, float(self.lb), float(self.ub))
            best_x = x1 if K["kernel"] == ["iso_gaussian"] else np.array([x1])
            self.train_stats = {"evals": obj.nevals, "batches": obj.nbatches}
            if disp:
                print(f"Best hyperparameter is {best_x}")
                print(f"With NegLnLikelihood of {neglnlikecand}")
        else:
            opt = K["optimizer"]
            if opt in ("lbfgsb", "slsqp"):
                optimbound = np.transpose(np.vstack((K["lbhyp"], K["ubhyp"])))
            elif opt == "cobyla":
                optimbound = []
                for i in range(len(K["ubhyp"])):
                    optimbound.append(lambda x, a=None, b=None, c=None, it=i: x[it] - K["lbhyp"][it])
                    optimbound.append(lambda x, a=None, b=None, c=None, it=i: K["ubhyp"][it] - x[it])
            else:
                optimbound = None
            if disp:
                print(f"Training {K['nrestart']} hyperparameter(s)")
            bestxcand, neglnlikecand = self.parallelopt(xhyp, n_cpu, optimbound, disp=disp, pool=pool, seed=seed)
            I = int(np.argmin(neglnlikecand))
            best_x = bestxcand[I, :]
            if disp:
                print("Single Objective, train hyperparam, end.")
                print(f"Best hyperparameter is {best_x}")
                print(f"With NegLnLikelihood of {neglnlikecand[I]}")
        self.KrigInfo = likelihood(best_x, K, mode="all", trainvar=self.trainvar)

    def parallelopt(self, xhyp, n_cpu, optimbound, disp=True, pool=None, seed=None):
        """Optimise from every start point (kriging_model.py:327-392); returns
        (bestxcand[nrestart, nbhyp], neglnlikecand[nrestart])."""
        K = self.KrigInfo
        xhyp = np.atleast_2d(np.asarray(xhyp, dtype=float))[:max(1, K["nrestart"])]
        bestx, bestf, obj = _tune_many(K, xhyp, self.trainvar, K["ubhyp"], K["lbhyp"], self.sigmacmaes,
                                       self.scaling, optimbound, seed=seed)
        self.train_stats = {"evals": obj.nevals, "batches": obj.nbatches}
        if disp:
            print(f"{obj.nevals} likelihood evaluations in {obj.nbatches} GPU batches")
        return bestx, bestf

    def loocvcalc(self, metrictype="mape", type="standard"):
        """Leave-one-out cross validation (kriging_model.py:260-297).  'standard'
        (krigloocv2.py: n refits in the reference) is one closed-form GPU pass over the factor;
        'fast' (krigloocv.py: the bordered inverse, Psi without nugget) is a closed form over a second factor
        (see :mod:`kadal_b200.loocv`)."""
        if type not in ("fast", "standard"):
            raise ValueError('"type" only accept fast or standard')
        from .loocv import loocv
        self.KrigInfo["LOOCVerror"], self.KrigInfo["LOOCVpred"] = loocv(self.KrigInfo, errtype=metrictype, type=type)
        return self.KrigInfo["LOOCVerror"], self.KrigInfo["LOOCVpred"]

    def predict(self, x, predtypes=["pred"], drmmodel=None):
        """Kriging prediction at raw sites x (kriging_model.py:299-325)."""
        return prediction(x, self.KrigInfo, predtypes=predtypes, drm=drmmodel)


def kriginfocheck(KrigInfo, lb, ub, nbhyp):
    """Validate KrigInfo and fill defaults (kriging_model.py:461-574): nrestart=1,
    optimizer='lbfgsb', TrendOrder=0, nugget=-6 fixed (or tuned for a [lo, hi] list), kernel
    list + nkernel, n_princomp, limittype, and the hyper-parameter box lbhyp / ubhyp with the
    CMA-ES scaling vector."""
    eps = np.finfo(float).eps
    if "nrestart" not in KrigInfo:
        KrigInfo["nrestart"] = 1
    elif KrigInfo["nrestart"] < 1:
        raise ValueError('Minimum value of KrigInfo["nrestart"] is 1')
    if "optimizer" not in KrigInfo:
        KrigInfo["optimizer"] = "lbfgsb"
        print("The optimizer is not specified, set to lbfgsb")
    elif KrigInfo["optimizer"].lower() not in ["lbfgsb", "cmaes", "cobyla", "slsqp", "ga"]:
        raise ValueError(KrigInfo["optimizer"], " is not a valid optimizer.")
    if "TrendOrder" not in KrigInfo:
        KrigInfo["TrendOrder"] = 0
    elif KrigInfo["TrendOrder"] < 0:
        raise ValueError("The order of the polynomial trend should be a positive value.")

    nnugget = 1
    if "nugget" not in KrigInfo:
        KrigInfo["nugget"] = -6
        KrigInfo["nuggetparam"] = "fixed"
        print("Nugget is not defined, set nugget to 1e-06")
    elif type(KrigInfo["nugget"]) is not list:
        KrigInfo["nuggetparam"] = "fixed"
    elif len(KrigInfo["nugget"]) == 2:
        KrigInfo["nuggetparam"] = "tuned"
        nnugget = 2
        if KrigInfo["nugget"][0] > KrigInfo["nugget"][1]:
            raise TypeError("The lower bound of the nugget should be lower than the upper bound.")
    else:
        raise TypeError("KrigInfo['nugget'] must be a number (fixed) or a [lower, upper] list (tuned)")

    if "kernel" not in KrigInfo:
        KrigInfo["kernel"] = ["gaussian"]
        print("Kernel is not defined, set kernel to gaussian")
    elif type(KrigInfo["kernel"]) is not list:
        KrigInfo["kernel"] = [KrigInfo["kernel"]]
    nkernel = len(KrigInfo["kernel"])
    KrigInfo["nkernel"] = nkernel

    if KrigInfo["type"] == "kriging":
        KrigInfo["n_princomp"] = False
    elif not ("n_princomp" in KrigInfo and KrigInfo["type"] == "kpls"):
        raise ValueError("KrigInfo['type'] = %r needs 'n_princomp' (kpls) or must be 'kriging'" % KrigInfo["type"])

    if "limit" in KrigInfo and "limittype" not in KrigInfo:
        KrigInfo["limittype"] = ">="
        print("KrigInfo['limittype'] is set to greater than equal'>='.")

    lbhyp = lb * np.ones(shape=[nbhyp])
    ubhyp = ub * np.ones(shape=[nbhyp])
    if nnugget > 1:
        lbhyp = np.hstack((lbhyp, KrigInfo["nugget"][0]))
        ubhyp = np.hstack((ubhyp, KrigInfo["nugget"][1]))
    if nkernel > 1:
        lbhyp = np.hstack((lbhyp, np.zeros(shape=[nkernel]) + eps))
        ubhyp = np.hstack((ubhyp, np.ones(shape=[nkernel])))
    scaling = np.ones(len(lbhyp))
    if nnugget > 1:
        scaling[nbhyp] = (KrigInfo["nugget"][1] - KrigInfo["nugget"][0]) / (ub - lb)
    if nkernel > 1:
        scaling[-nkernel:] = 1 / (ub - lb)
    KrigInfo["lbhyp"] = lbhyp
    KrigInfo["ubhyp"] = ubhyp
    return KrigInfo, scaling
