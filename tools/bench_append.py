"""One-sample append (kb200_fit_append) against a full device refit (kb200_fit_state) at the same hyper-parameters:
wall time of the blocking C-ABI call, with and without the n x n U / Psi copies back to the host.
    python tools/bench_append.py [n] [d] > gpurun_out/bench_append.json"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kadal_b200.model import GpuModel   # noqa: E402


def run(n, d, reps=8):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n + reps, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n + reps) + 3.0
    x = np.full(d, 0.3)
    res = {"n": n, "d": d, "reps": reps}
    mk = lambda k: GpuModel(X[:n + k], y[:n + k], np.ones((n + k, 1)), [0])
    for want in (False, True):
        prev = mk(0)
        prev.fit_state(x, False, -6.0, want_U=want, want_Psi=want)
        cold_app, cold_full, warm_app, warm_full, dn = [], [], [], [], []
        for k in range(1, reps + 1):
            # cold: model creation + first call, as a believer / AK-MCS / SOBO update sees it
            t0 = time.perf_counter()
            nxt = mk(k)
            a = nxt.fit_state(x, False, -6.0, want_U=want, want_Psi=want, append_to=prev)
            t1 = time.perf_counter()
            ref = mk(k)
            b = ref.fit_state(x, False, -6.0, want_U=want, want_Psi=want)
            t2 = time.perf_counter()
            cold_app.append((t1 - t0) * 1e3)
            cold_full.append((t2 - t1) * 1e3)
            # warm: the same calls again (workspaces allocated)
            t0 = time.perf_counter()
            nxt.fit_state(x, False, -6.0, want_U=want, want_Psi=want, append_to=prev)
            t1 = time.perf_counter()
            ref.fit_state(x, False, -6.0, want_U=want, want_Psi=want)
            t2 = time.perf_counter()
            warm_app.append((t1 - t0) * 1e3)
            warm_full.append((t2 - t1) * 1e3)
            dn.append(abs(a["nll"] - b["nll"]) / max(1.0, abs(b["nll"])))
            prev.close()
            ref.close()
            prev = nxt
        key = "with_U_Psi" if want else "device_only"
        med = lambda v: float(np.median(v))
        res[key] = {"cold_append_ms": med(cold_app), "cold_refit_ms": med(cold_full), "warm_append_ms": med(warm_app),
                    "warm_refit_ms": med(warm_full), "nll_rel_diff_max": float(max(dn)), "nll_last": [a["nll"], b["nll"]], "info_last": [a["info"], b["info"]]}
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    out = [run(n, d)]
    if len(sys.argv) <= 1:
        out.append(run(4000, 20, reps=4))
    print(json.dumps(out))
