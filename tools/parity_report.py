"""GPU-vs-oracle parity report with condition numbers (writes gpurun_out/parity_report.json).

Test infrastructure: uses the oracle as the checker.  For every candidate it records
cond(Psi + eps I), the oracle NLL and the CUDA NLL, so that the error can be read against
the fp64 noise floor cond * ulp."""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import kriging_oracle as ko  # noqa: E402
import kadal_b200 as kb  # noqa: E402
import cases  # noqa: E402
from conftest import load_golden  # noqa: E402


def report(name, info, pop, nmax):
    nll, inf = kb.likelihood_batch(pop, info, False)
    rows = []
    for i in range(min(nmax, len(pop))):
        A = copy.deepcopy(info)
        ko.likelihood(pop[i].copy(), A, "all", False, check_pd=False)
        n = A["Psi"].shape[0]
        cond = float(np.linalg.cond(A["Psi"] + np.eye(n) * 10.0 ** float(A["nugget"])))
        ref = float(A["NegLnLike"])
        rows.append({"i": i, "cond": cond, "ref": ref, "gpu": float(nll[i]),
                     "rel": abs(float(nll[i]) - ref) / max(1.0, abs(ref))})
    worst = max(rows, key=lambda r: r["rel"])
    print("%s: %d candidates, worst rel %.2e at cond %.2e; max rel/(cond*2^-53) = %.3f" % (
        name, len(rows), worst["rel"], worst["cond"],
        max(r["rel"] / (r["cond"] * 2.0 ** -53) for r in rows)), flush=True)
    return rows


def main():
    out = {}
    g = load_golden("c3_like.npz")
    info, _ = cases.c3_info(g)
    out["c3_pop"] = report("c3_pop", info, g["pop"], 48)
    w = ko.workload("C2")
    info2 = ko.make_kriginfo(w["X"], w["y"])
    out["c2_popB"] = report("c2_popB", info2, w["popB"], 24)
    out["c2_popA"] = report("c2_popA", info2, w["popA"], 8)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
