"""BASELINE config 3 at full size as an AK-MCS iteration sees it: the learning-function / stopping reductions of
10^7 Monte-Carlo points held in an ordinary (pageable) numpy array, one call (reliability.akmcs_reductions).
    python tools/bench_akmcs_full.py [npts]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import kadal_b200 as kb                          # noqa: E402
from kadal_b200 import reliability              # noqa: E402
from oracle import kriging_oracle as ko         # noqa: E402  (make_kriginfo only: no oracle arithmetic is timed)

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(1)
n, d = 200, 6
X = rng.uniform(-1, 1, (n, d))
y = (np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2)[:, None]
info = ko.make_kriginfo(X, y)
info["limit"], info["limittype"] = 0.0, ">="
kb.likelihood(np.full(d, 0.2), info, "all", False)
pop = np.clip(np.random.default_rng(2).standard_normal((npts, d)) * 0.5, -1, 1)
ts, res = [], None
for rep in range(4):
    t0 = time.perf_counter()
    res = reliability.akmcs_reductions(pop, info)
    ts.append(time.perf_counter() - t0)
best = min(ts[1:])
print(json.dumps({"workload": "C3 AK-MCS iteration: n=200, d=6, U-function reductions over %d pageable MC points" % npts,
                  "ms": best * 1e3, "pts_per_s": npts / best, "first_call_ms": ts[0] * 1e3,
                  "Pf": res["Pf"], "minU": res["minU"], "minUloc": res["minUloc"]}))
