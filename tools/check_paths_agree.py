import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from kadal_b200.model import GpuModel
rng = np.random.default_rng(5)
for n, d in ((2048, 8), (2100, 8), (1537, 12), (385, 3)):
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1))[:, None] + 0.1 * rng.standard_normal((n, 1))
    pop = rng.uniform(-0.3, 1.0, (12, d))
    out = {}
    for h in ("0", "1"):
        os.environ["KB200_HALF"] = h
        m = GpuModel(X, y, np.ones((n, 1)), [0, 2])   # composite gaussian + matern32, fixed weights via x
        pp = np.hstack([pop, rng.uniform(0.2, 1, (12, 2))]) if h == "0" else pp
        nll, inf = m.lik_batch(pp, False, -6.0)
        nll2, _ = m.lik_batch(pp, False, -6.0)
        out[h] = nll
        assert np.array_equal(nll, nll2) and not inf.any(), (n, h, inf)
        m.close()
    print(n, d, "max rel diff half vs full", float(np.max(np.abs(out["0"] - out["1"]) / np.maximum(1, np.abs(out["0"])))), "argmin", int(np.argmin(out["0"])), int(np.argmin(out["1"])))
