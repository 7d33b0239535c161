import os, sys, copy, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from kadal_b200.model import GpuModel
from oracle import kriging_oracle as ko
X, y, pop = bench.c2_data()
n = 1000
res = {}
for name, env in (("full", {"KB200_HALF": "0"}), ("left", {"KB200_HALF": "1", "KB200_RL": "0"}), ("right", {"KB200_HALF": "1", "KB200_RL": "1"})):
    for k in ("KB200_HALF", "KB200_RL"): os.environ.pop(k, None)
    os.environ.update(env)
    m = GpuModel(X, y, np.ones((n, 1)), [0]); res[name], _ = m.lik_batch(pop, False, -6.0); m.close()
d = np.abs(res["left"] - res["full"]) / np.maximum(1, np.abs(res["full"]))
idx = np.argsort(d)[-3:]
info = ko.make_kriginfo(X, y)
for i in idx:
    A = copy.deepcopy(info)
    ref = ko.likelihood(pop[i].copy(), A, "all", False, check_pd=False)["NegLnLike"]
    cond = np.linalg.cond(A["Psi"] + np.eye(n) * 1e-6)
    print("cand", i, "cond %.2e" % cond, "tol %.2e" % max(1e-9, 0.5 * cond * 2.0 ** -53),
          {k: "%.2e" % (abs(v[i] - ref) / max(1, abs(ref))) for k, v in res.items()})
