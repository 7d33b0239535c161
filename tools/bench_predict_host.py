"""prediction() as a KADAL caller issues it: ordinary (pageable) numpy input, results as numpy arrays.
Wall time of GpuModel.predict for C3 (n=200, d=6, pred + s) and C4 (n=800, d=20, pred) point sets.
    python tools/bench_predict_host.py            # results in recycled pinned buffers (default)
    KB200_PINNED_MB=0 python tools/bench_predict_host.py   # pageable results"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kadal_b200.model import GpuModel   # noqa: E402

out = {"pinned_mb": os.environ.get("KB200_PINNED_MB", "2048")}
for name, n, d, npts, want in (("c3", 200, 6, 4_000_000, 1 | 4), ("c4", 800, 20, 2_000_000, 1)):
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2 + 3.0
    m = GpuModel(X, y, np.ones((n, 1)), [0])
    m.fit_state(np.full(d, 0.2), False, -6.0, want_U=False, want_Psi=False)
    x = rng.uniform(-1, 1, (npts, d))
    lb, ub = -np.ones(d), np.ones(d)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        r = m.predict(x, want, normtype=1, norm_a=lb, norm_b=ub)
        ts.append(time.perf_counter() - t0)
        chk = float(r["pred"][::1000].sum())
        del r
    out[name] = {"npts": npts, "ms_best": min(ts[1:]) * 1e3, "pts_per_s": npts / min(ts[1:]), "check": chk}
    # the same with the points already in pinned memory (what bench.py's e2e measures)
    from kadal_b200 import _capi
    xp = _capi.pinned_empty(x.shape)
    xp[:] = x
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        r = m.predict(xp, want, normtype=1, norm_a=lb, norm_b=ub)
        ts.append(time.perf_counter() - t0)
        del r
    out[name]["pinned_input_ms_best"] = min(ts[1:]) * 1e3
    _capi.pinned_free(xp)
    m.close()
print(json.dumps(out))
