"""First-contact GPU diagnostics: kernel-by-kernel parity against the oracle (test infra)."""
import copy
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kriging_oracle as ko  # noqa: E402
import kadal_b200 as kb  # noqa: E402
from kadal_b200.model import GpuModel, model_for, clear_cache  # noqa: E402


def rel(a, b, floor=1.0):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def case(n, d, kernel=("gaussian",), seed=0, pls=None, npc=False):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(2 * X.sum(1, keepdims=True)) + X[:, :1] ** 2 + 3.0
    return ko.make_kriginfo(X, y, kernel=kernel, pls=pls, n_princomp=npc)


def check_case(name, info, x, trainvar=False, npred=300):
    A = copy.deepcopy(info)
    B = copy.deepcopy(info)
    ko.likelihood(np.array(x, dtype=float), A, "all", trainvar, check_pd=False)
    m = model_for(B)
    psi = m.debug_psi(x, trainvar, B["nugget"] if not isinstance(B["nugget"], list) else B["nugget"][0])
    eq = np.mean(psi == A["Psi"])
    out = kb.likelihood(np.array(x, dtype=float), B, "all", trainvar)
    xp = np.random.default_rng(99).uniform(-1, 1, (npred, info["nvar"]))
    pa = ko.prediction(xp, A, ["pred", "SSqr", "s"])
    pb = kb.prediction(xp, B, ["pred", "SSqr", "s"])
    # prediction from the oracle's own U (predictor_load path)
    C = copy.deepcopy(A)
    pc = kb.prediction(xp, C, ["pred", "SSqr"])
    s2 = float(np.asarray(A["SigmaSqr"]).ravel()[0])
    print("%-28s n=%4d psi_eq=%.4f psi_rel=%.1e nll=%.10g d_nll=%.1e U=%.1e BE=%.1e s2=%.1e | pred=%.1e ssqr=%.1e | loadU pred=%.1e ssqr=%.1e"
          % (name, info["nsamp"], eq, rel(psi, A["Psi"], 1e-300), A["NegLnLike"],
             abs(B["NegLnLike"] - A["NegLnLike"]) / max(1, abs(A["NegLnLike"])),
             rel(B["U"], A["U"]), rel(B["BE"], A["BE"]), rel(B["SigmaSqr"], A["SigmaSqr"], 1e-300),
             rel(pb[0], pa[0]), rel(pb[1], pa[1], s2 * 1e-6), rel(pc[0], pa[0]), rel(pc[1], pa[1], s2 * 1e-6)), flush=True)


def main():
    print("naive =", os.environ.get("KB200_NAIVE", "0"))
    x4 = [0.1, -0.2, 0.3, 0.0]
    check_case("gauss n64", case(64, 4), x4)
    check_case("gauss n128", case(128, 4), x4)
    check_case("gauss n200", case(200, 4), x4)
    check_case("gauss n300", case(300, 4), x4)
    check_case("expo n200", case(200, 4, ("exponential",)), x4)
    check_case("mat32 n200", case(200, 4, ("matern32",)), x4)
    check_case("mat52 n200", case(200, 4, ("matern52",)), x4)
    check_case("trainvar n200", case(200, 4), x4 + [0.5], trainvar=True)
    check_case("tuned n200", case(200, 4), x4 + [-4.0])
    check_case("composite n200", case(200, 4, ("gaussian", "matern32", "exponential")), x4 + [.5, .3, .2])
    pls = np.random.default_rng(3).standard_normal((4, 2))
    check_case("kpls n200", case(200, 4, pls=pls, npc=2), [0.1, -0.2])
    check_case("gauss d10 n500", case(500, 10), list(np.linspace(-0.2, 0.6, 10)))
    check_case("gauss d20 n300", case(300, 20), list(np.linspace(0.2, 0.8, 20)))
    # C2 KAT
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    t = time.time()
    v = kb.likelihood(w["kat_x"], info, "default", False)
    print("C2 KAT nll=%.13f (ref -345.9939996418981) rel=%.2e  first-call %.3fs" %
          (v, abs(v + 345.9939996418981) / 345.99, time.time() - t), flush=True)
    for pop in ("popA", "popB"):
        t = time.time()
        nll, inf = kb.likelihood_batch(w[pop], info, False)
        t1 = time.time() - t
        t = time.time()
        nll, inf = kb.likelihood_batch(w[pop], info, False)
        t2 = time.time() - t
        ref = ko.likelihood_population(w[pop][:6], ko.make_kriginfo(w["X"], w["y"]), False, check_pd=False)
        print("C2 %s: 512 cand %.3fs / %.3fs -> %.0f evals/s; info!=0: %d; first6 rel=%.2e argmin=%d" %
              (pop, t1, t2, 512 / t2, int((inf != 0).sum()), rel(nll[:6], ref), int(np.argmin(nll))), flush=True)


if __name__ == "__main__":
    main()
