"""pred + s on small models (n <= 256): the fused predictor (predict_small_kernel) against the general tiled path
(psi kernel + pure-solve / panel kernels), 4e6 points.  Decides the policy in predict_core."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_extra as be
from kadal_b200 import _capi
want = _capi.OUT_PRED | _capi.OUT_S
for n, d in ((40, 2), (64, 4), (100, 6), (128, 6), (129, 6), (160, 6), (200, 6), (200, 8), (232, 6), (256, 6)):
    rng = np.random.default_rng(n + d)
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2 + 3.0
    row = {}
    for name, env in (("fused", "1"), ("general", None)):
        os.environ.pop("KB200_FUSED_PREDICT", None)
        if env: os.environ["KB200_FUSED_PREDICT"] = env
        r = be.pred_case(torch, name, X, y, np.full(d, 0.2), 4_000_000, want)
        row[name] = r["pts_per_s"]
    print("n %3d d %2d  fused %.3e  general %.3e  ratio %.2f" % (n, d, row["fused"], row["general"], row["general"] / row["fused"]), flush=True)
