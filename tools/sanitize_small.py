"""Small pass over every kernel family for compute-sanitizer (memcheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kadal_b200 as kb
from kadal_b200 import ehvi, loocv
from oracle import kriging_oracle as ko
rng = np.random.default_rng(0)
for n, d, kern in ((300, 5, ("gaussian",)), (130, 3, ("matern52", "exponential")), (385, 4, ("gaussian",))):
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1, keepdims=True)) + 3.0
    info = ko.make_kriginfo(X, y, kernel=kern)
    nb = d + (len(kern) if len(kern) > 1 else 0)
    pop = rng.uniform(-0.5, 0.8, (40, nb)); pop[:, d:] = rng.uniform(0.2, 1, (40, nb - d))
    for P in (40, 3):
        nll, inf = kb.likelihood_batch(pop[:P], info, False)      # left- and right-looking schedules
        assert not inf.any()
    kb.likelihood(pop[0].copy(), info, "all", False)
    out = kb.prediction(rng.uniform(-1, 1, (700, d)), info, ["pred", "SSqr", "s", "EI", "PoI"])
    loo = loocv.loocv(info) if hasattr(loocv, "loocv") else None
    print("n", n, kern, "ok", float(nll[0]), float(out[0][0, 0]), flush=True)
X = rng.uniform(-1, 1, (260, 12)); y = (X[:, :3].sum(1) ** 2)[:, None] + 3.0
info = ko.make_kriginfo(X, y, pls=rng.standard_normal((12, 2)) / 4, n_princomp=2)
nll, inf = kb.likelihood_batch(rng.uniform(-0.3, 0.5, (12, 2)), info, False)
print("kpls ok", float(nll[0]), flush=True)
v = np.abs(rng.standard_normal((15, 3))); P3 = -10 + 9.5 * v / np.linalg.norm(v, axis=1, keepdims=True)
for mode in ("reference", "independent"):
    h = ehvi.ehvi3d_sliceupdate_multi(P3, [-10.0] * 3, rng.uniform(-10, -1, (70, 3)), rng.uniform(0.1, 2, (70, 3)), mode=mode)
a, b = np.sort(rng.uniform(0, 1, 9)), np.sort(rng.uniform(0, 1, 9))[::-1]
h2 = ehvi.exi2d_batch(np.c_[a, b], [1.2, 1.3], rng.uniform(0, 1.2, (50, 2)), rng.uniform(0.01, 0.4, (50, 2)))
print("ehvi ok", float(h.max()), float(h2.max()), flush=True)
