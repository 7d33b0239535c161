"""Extended-precision (x87 long double, 64-bit significand) ln-likelihood of the 12 worst-conditioned
candidates of C2 population B (and candidate 114), from the SAME float64 Psi the reference builds:

    OMP_NUM_THREADS=2 PYTHONPATH=/root/repo python tests/golden/make_truth.py

Writes tests/golden/c2_truth.npz (idx, truth, ref).  It shows how far the reference's own float64 result
is from the exact value of its formula (|ref - truth| ~ cond * ulp): the scale any parity tolerance at
cond ~ 1e8..1e9 has to be read against."""
import copy, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import kriging_oracle as ko  # noqa: E402
LD = np.longdouble


def nll_longdouble(Psi, y, eps=1e-6):
    n = Psi.shape[0]
    A = (Psi + np.eye(n) * eps).astype(LD)          # the reference adds eps in float64 (likelihood_func.py:171)
    L = np.zeros((n, n), dtype=LD)
    for k in range(n):
        d = np.sqrt(A[k, k]); L[k, k] = d
        col = A[k + 1:, k] / d; L[k + 1:, k] = col
        A[k + 1:, k + 1:] -= np.outer(col, col)
    logdet = 2 * np.sum(np.log(np.diag(L)))

    def fs(b):
        x = b.astype(LD).copy()
        for k in range(n):
            x[k] /= L[k, k]; x[k + 1:] -= L[k + 1:, k] * x[k]
        return x
    z1, zy = fs(np.ones(n)), fs(y.ravel())
    mu = (z1 @ zy) / (z1 @ z1)
    r = zy - mu * z1
    s2 = (r @ r) / n
    return float((n / LD(2)) * np.log(s2) + LD(0.5) * logdet)      # likelihood_func.py:202-204


def one(i):
    w = ko.workload("C2"); info = ko.make_kriginfo(w["X"], w["y"])
    ref = ko.likelihood(w["popB"][i].copy(), info, "all", False, check_pd=False)["NegLnLike"]
    return i, nll_longdouble(info["Psi"], w["y"]), ref


if __name__ == "__main__":
    from multiprocessing import get_context
    g = np.load(os.path.join(HERE, "c2_full.npz"))
    # the 12 worst-conditioned + candidate 114 (|NLL| = 0.07, the largest GPU-vs-reference difference of the population)
    idx = np.unique(np.r_[np.argsort(g["cond_popB"])[-12:], 114])
    with get_context("spawn").Pool(4) as pool:
        res = sorted(pool.map(one, [int(i) for i in idx]))
    out = dict(idx=np.array([r[0] for r in res]), truth=np.array([r[1] for r in res]), ref=np.array([r[2] for r in res]),
               cond=g["cond_popB"][[r[0] for r in res]])
    np.savez_compressed(os.path.join(HERE, "c2_truth.npz"), **out)
    for r, c in zip(res, out["cond"]):
        print(r[0], "cond %.2e" % c, "ref - truth %.2e" % (r[2] - r[1]), "in cond*ulp: %.2f" % (abs(r[2] - r[1]) / max(1, abs(r[1])) / (c * 2.0 ** -53)))
