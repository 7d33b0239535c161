"""Small, fixed workloads for ncu captures (see profiles/README.md).

    python tools/prof_run.py lik  [passes]     # C2: n=1000, d=10, 512 candidates
    python tools/prof_run.py pred [passes]     # C3: n=200, d=6, pred + s over 2^20 points
    python tools/prof_run.py pred4 [passes]    # C4: n=800, d=20, pred over 2^20 points
    python tools/prof_run.py small [passes]    # n=128, d=6, 512 candidates: lik_small_kernel (one CTA per candidate in shared memory)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kadal_b200 import _capi  # noqa: E402
from kadal_b200.model import GpuModel  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "lik"
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    if mode == "lik":
        X, y, pop = bench.c2_data()
        m = GpuModel(X, y, np.ones((len(X), 1)), [0])
        for _ in range(passes):
            nll, info = m.lik_batch(pop, False, -6.0)
        print("lik ok", float(nll.min()), int((info != 0).sum()))
    elif mode == "small":
        rng = np.random.default_rng(23)
        X = rng.uniform(-1, 1, (128, 6))
        y = (np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2)[:, None] + 3.0
        m = GpuModel(X, y, np.ones((128, 1)), [0])
        pop = rng.uniform(-0.5, 1.0, (512, 6))
        for _ in range(passes):
            nll, info = m.lik_batch(pop, False, -6.0)
        print("small ok", float(nll.min()), int((info != 0).sum()))
    else:
        if mode == "pred":
            X, y, th = bench.c3_data()
            want = _capi.OUT_PRED | _capi.OUT_S
        else:
            rng = np.random.default_rng(7)
            X = rng.uniform(-1, 1, (800, 20))
            y = np.sin(X[:, :3].sum(1))[:, None] + 0.05 * rng.standard_normal((800, 1)) + 3.0
            th = np.full(20, 0.5)
            want = _capi.OUT_PRED
        n, d = X.shape
        m = GpuModel(X, y, np.ones((n, 1)), [0])
        st = m.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
        assert st["info"] == 0
        xs = np.clip(np.random.default_rng(2).standard_normal((1 << 20, d)) * 0.5, -1, 1)
        for _ in range(passes):
            out = m.predict(xs, want, _capi.NORM_DEFAULT, -np.ones(d), np.ones(d))
        print("pred ok", float(out["pred"].mean()))


if __name__ == "__main__":
    main()
