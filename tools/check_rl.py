"""Right-looking variant (KB200_RL=1) against the default path on the same populations."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from kadal_b200.model import GpuModel
rng = np.random.default_rng(5)
for n, d in ((700, 8), (1000, 10), (2300, 6)):
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1))[:, None] + 0.1 * rng.standard_normal((n, 1))
    pop = rng.uniform(-0.3, 1.0, (20, d))
    out = {}
    for rl in ("0", "1"):
        os.environ["KB200_RL"] = rl
        m = GpuModel(X, y, np.ones((n, 1)), [0])
        nll, inf = m.lik_batch(pop, False, -6.0)
        nll2, _ = m.lik_batch(pop, False, -6.0)
        st = m.fit_state(pop[0], False, -6.0, want_Psi=True)
        A = st["Psi"] + np.eye(n) * 1e-6
        res = np.linalg.norm(st["U"].T @ st["U"] - A) / np.linalg.norm(A)
        out[rl] = nll
        assert np.array_equal(nll, nll2) and not inf.any(), (n, rl, inf)
        print(n, "rl", rl, "factor residual %.2e" % res, "beta", st["BE"], "nll0", nll[0])
        m.close()
    print(n, d, "max rel diff", float(np.max(np.abs(out["0"] - out["1"]) / np.maximum(1, np.abs(out["0"])))))
