"""Every candidate of both C2 populations against the live-reference goldens (tests/golden/c2_full.npz):
distribution of |gpu - ref| / max(1, |ref|) in units of cond(Psi + eps I) * 2^-53, per factorisation schedule."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kriging_oracle as ko
from kadal_b200.model import GpuModel
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c2_full.npz"))
w = ko.workload("C2")
n = 1000
for name, env in (("left-looking (population)", {"KB200_RL": "0"}), ("right-looking", {"KB200_RL": "1"}), ("full-SM legacy", {"KB200_HALF": "0"})):
    for k in ("KB200_HALF", "KB200_RL"): os.environ.pop(k, None)
    os.environ.update(env)
    m = GpuModel(w["X"], w["y"], np.ones((n, 1)), [0])
    for pop in ("popA", "popB"):
        nll, inf = m.lik_batch(w[pop], False, -6.0)
        ref, cond = g["nll_" + pop], g["cond_" + pop]
        rel = np.abs(nll - ref) / np.maximum(1.0, np.abs(ref))
        u = rel / (cond * 2.0 ** -53)
        print("%-26s %s: info!=0 %d | max rel %.2e (cond %.1e) | rel/(cond ulp): median %.3f p99 %.3f max %.3f | argmin %d vs %d | > tol: %d" % (
            name, pop, int((inf != 0).sum()), rel.max(), cond[rel.argmax()], np.median(u), np.quantile(u, 0.99), u.max(),
            int(np.argmin(nll)), int(np.argmin(ref)), int((rel > np.maximum(1e-9, 0.5 * cond * 2.0 ** -53)).sum())), flush=True)
    m.close()
# the 12 worst-conditioned popB candidates against the extended-precision value of the same formula
tp = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c2_truth.npz")
if os.path.exists(tp):
    t = np.load(tp)
    for k in ("KB200_HALF", "KB200_RL"): os.environ.pop(k, None)
    m = GpuModel(w["X"], w["y"], np.ones((n, 1)), [0])
    nll, _ = m.lik_batch(w["popB"], False, -6.0)
    m.close()
    den = np.maximum(1.0, np.abs(t["truth"])) * t["cond"] * 2.0 ** -53
    eg, er = np.abs(nll[t["idx"]] - t["truth"]) / den, np.abs(t["ref"] - t["truth"]) / den
    print("vs long-double truth, in cond*ulp: gpu median %.3f max %.3f | reference median %.3f max %.3f" % (
        np.median(eg), eg.max(), np.median(er), er.max()))
    for i, a, b in zip(t["idx"], eg, er): print("  cand %3d gpu %.3f ref %.3f" % (i, a, b))
