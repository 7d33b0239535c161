"""Ask/tell (mu/mu_w, lambda)-CMA-ES for the hyper-parameter search of ``Kriging.train``.

The reference calls ``cma.fmin2(likelihood, x0, sigma0, {'bounds', 'scaling_of_variables'})``
(kriging_model.py:423-428), which evaluates ONE candidate per objective call.  Here the
whole population of a generation — and the populations of all restarts, which advance in
lock step — is handed to one batched likelihood call on the GPU (``kb200_lik_batch``).

This is an independent implementation of the published algorithm (N. Hansen, "The CMA
Evolution Strategy: A Tutorial", 2016: default strategy parameters, rank-one + rank-mu
update, cumulative step-size adaptation); it is host-side control logic, O(lambda n^2) per
generation, next to O(lambda n_samp^3) on the device.  Box bounds are handled by projecting
the sampled point onto the box for evaluation and adding a quadratic penalty on the
projection distance, so the returned optimum is always feasible.
"""
import math

import numpy as np


class CMAES:
    """One CMA-ES run (minimisation).  Typical loop::

        es = CMAES(x0, sigma0, bounds=(lb, ub))
        while not es.stop():
            X = es.ask()                 # (popsize, n) feasible candidates
            es.tell(X, f(X))             # f evaluated for the whole population at once
        es.best_x, es.best_f
    """

    def __init__(self, x0, sigma0, bounds=None, scaling=None, popsize=None, seed=None,
                 maxiter=None, tolfun=1e-11, tolfunhist=1e-12, tolx=1e-11):
        self.n = n = int(np.size(x0))
        self.scale = np.ones(n) if scaling is None else np.asarray(scaling, dtype=float).reshape(n)
        if np.any(self.scale <= 0):
            raise ValueError("scaling_of_variables must be positive")
        self.lb = self.ub = None
        if bounds is not None:
            self.lb = np.broadcast_to(np.asarray(bounds[0], dtype=float), (n,)).copy()
            self.ub = np.broadcast_to(np.asarray(bounds[1], dtype=float), (n,)).copy()
            if np.any(self.ub < self.lb):
                raise ValueError("upper bound below lower bound")
        self.mean = self._project(np.asarray(x0, dtype=float).reshape(n).copy())
        self.sigma = float(sigma0)
        self.sigma0 = float(sigma0)
        self.lam = lam = int(popsize) if popsize else 4 + int(3 * math.log(n))
        self.mu = mu = lam // 2
        w = math.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
        self.w = w / w.sum()
        self.mueff = 1.0 / np.sum(self.w ** 2)
        me = self.mueff
        self.cc = (4 + me / n) / (n + 4 + 2 * me / n)
        self.cs = (me + 2) / (n + me + 5)
        self.c1 = 2 / ((n + 1.3) ** 2 + me)
        self.cmu = min(1 - self.c1, 2 * (me - 2 + 1 / me) / ((n + 2) ** 2 + me))
        self.damps = 1 + 2 * max(0.0, math.sqrt((me - 1) / (n + 1)) - 1) + self.cs
        self.chin = math.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        self.pc = np.zeros(n)
        self.ps = np.zeros(n)
        self.C = np.eye(n)
        self.B = np.eye(n)
        self.D = np.ones(n)
        self.invsqrtC = np.eye(n)
        self.eigen_iter = 0
        self.iter = 0
        self.evals = 0
        self.rng = np.random.default_rng(seed)
        self.maxiter = int(maxiter) if maxiter else int(100 + 150 * (n + 3) ** 2 / math.sqrt(lam))
        self.tolfun, self.tolfunhist, self.tolx = tolfun, tolfunhist, tolx
        self.hist = []
        self.best_x = self.mean.copy()
        self.best_f = np.inf
        self._y = None
        self._reason = None

    # ------------------------------------------------------------------ helpers
    def _project(self, x):
        if self.lb is None:
            return x
        return np.minimum(np.maximum(x, self.lb), self.ub)

    def _update_eigen(self):
        self.C = np.triu(self.C) + np.triu(self.C, 1).T
        vals, vecs = np.linalg.eigh(self.C)
        vals = np.maximum(vals, 1e-300)
        self.D = np.sqrt(vals)
        self.B = vecs
        self.invsqrtC = (vecs / self.D) @ vecs.T
        self.eigen_iter = self.iter

    # ------------------------------------------------------------------ ask / tell
    def ask(self):
        """Sample lambda candidates; returns the feasible (projected) points to evaluate."""
        if self.iter - self.eigen_iter > 1.0 / ((self.c1 + self.cmu) * self.n * 10):
            self._update_eigen()
        z = self.rng.standard_normal((self.lam, self.n))
        self._y = (z * self.D) @ self.B.T                      # ~ N(0, C)
        self._x = self.mean + self.sigma * self._y * self.scale
        self._xf = np.array([self._project(x) for x in self._x])
        return self._xf.copy()

    def tell(self, X, fvals):
        fvals = np.asarray(fvals, dtype=float).reshape(-1)
        if fvals.shape[0] != self.lam:
            raise ValueError("tell() needs one value per asked candidate")
        raw = np.where(np.isfinite(fvals), fvals, np.inf)
        self.evals += self.lam
        ib = int(np.argmin(raw))
        if raw[ib] < self.best_f:
            self.best_f = float(raw[ib])
            self.best_x = self._xf[ib].copy()
        # penalised ranking: infeasible samples pay for their distance to the box
        pen = np.sum(((self._x - self._xf) / self.scale) ** 2, axis=1)
        finite = raw[np.isfinite(raw)]
        spread = (np.percentile(finite, 75) - np.percentile(finite, 25)) if finite.size > 3 else 1.0
        spread = spread if spread > 0 else 1.0
        rank_f = np.where(np.isfinite(raw), raw, (finite.max() if finite.size else 0.0) + 1e3 * abs(spread)) \
            + pen * spread / max(self.sigma ** 2, 1e-300)
        order = np.argsort(rank_f, kind="stable")
        ysel = self._y[order[:self.mu]]
        yw = self.w @ ysel
        n = self.n
        self.mean = self.mean + self.sigma * yw * self.scale
        self.ps = (1 - self.cs) * self.ps + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * (self.invsqrtC @ yw)
        self.iter += 1
        hsig = (np.linalg.norm(self.ps) / math.sqrt(1 - (1 - self.cs) ** (2 * self.iter)) / self.chin) < (1.4 + 2 / (n + 1))
        self.pc = (1 - self.cc) * self.pc + (math.sqrt(self.cc * (2 - self.cc) * self.mueff) * yw if hsig else 0.0)
        rank_mu = (ysel.T * self.w) @ ysel
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (0.0 if hsig else self.cc * (2 - self.cc)) * self.C)
                  + self.cmu * rank_mu)
        self.sigma *= math.exp(min(1.0, (self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chin - 1)))
        self.hist.append(float(raw[order[0]]) if np.isfinite(raw[order[0]]) else float(rank_f[order[0]]))
        self._last_f = np.sort(rank_f)

    # ------------------------------------------------------------------ termination
    def stop(self):
        if self._reason:
            return self._reason
        if self.iter == 0:
            return None
        if self.iter >= self.maxiter:
            self._reason = "maxiter"
        elif not np.all(np.isfinite(self.C)) or not np.isfinite(self.sigma):
            self._reason = "numerical"
        else:
            h = self.hist[-(10 + int(math.ceil(30 * self.n / self.lam))):]
            rng_now = self._last_f[-1] - self._last_f[0] if np.all(np.isfinite(self._last_f)) else np.inf
            if len(h) >= 10 and (max(h) - min(h)) < self.tolfunhist:
                self._reason = "tolfunhist"
            elif len(h) >= 10 and max(max(h) - min(h), rng_now) < self.tolfun:
                self._reason = "tolfun"
            elif (np.all(self.sigma * np.sqrt(np.diag(self.C)) * self.scale < self.tolx)
                  and np.all(self.sigma * np.abs(self.pc) * self.scale < self.tolx)):
                self._reason = "tolx"
            elif self.D.max() / self.D.min() > 1e7:
                self._reason = "conditioncov"
        return self._reason


def fmin_lockstep(batch_fun, x0s, sigma0, bounds=None, scaling=None, popsize=None, seed=None,
                  maxiter=None, tolfun=1e-11, tolx=1e-11, callback=None):
    """Run one CMA-ES per row of ``x0s`` in lock step: every generation the populations of
    all still-running restarts are concatenated and evaluated by ONE call
    ``batch_fun(X[(k*lambda), n]) -> f[(k*lambda)]``.  Returns (best_x[k, n], best_f[k], runs)."""
    x0s = np.atleast_2d(np.asarray(x0s, dtype=float))
    ss = np.random.SeedSequence(seed)
    runs = [CMAES(x0, sigma0, bounds, scaling, popsize, seed=s, maxiter=maxiter, tolfun=tolfun, tolx=tolx)
            for x0, s in zip(x0s, ss.spawn(len(x0s)))]
    while True:
        live = [es for es in runs if not es.stop()]
        if not live:
            break
        pops = [es.ask() for es in live]
        f = np.asarray(batch_fun(np.vstack(pops)), dtype=float).reshape(-1)
        k = 0
        for es, pop in zip(live, pops):
            es.tell(pop, f[k:k + len(pop)])
            k += len(pop)
        if callback is not None:
            callback(runs)
    return (np.array([es.best_x for es in runs]), np.array([es.best_f for es in runs]), runs)
