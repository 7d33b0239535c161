"""A/B: prediction with d^2 * RN(1/theta^2) (the build) against d^2 / theta^2 exactly as numpy divides
(-DKB200_PREDICT_EXACT_DIV=1), both against the oracle, on the C3 and C4 models.

    python tools/ab_predict_div.py build      # build container: second library into kadal_b200/_lib/ab/
    python tools/ab_predict_div.py            # GPU box: runs both, prints error statistics
"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
AB = os.path.join(ROOT, "kadal_b200", "_lib", "ab", "libkb200_exactdiv.so")


def build():
    from kadal_b200 import _build
    os.makedirs(os.path.dirname(AB), exist_ok=True)
    objs = []
    for src in _build.SOURCES:
        obj = os.path.join(os.path.dirname(AB), src.replace(".cu", ".o"))
        subprocess.run([_build._nvcc()] + _build.NVCC_FLAGS + ["-DKB200_PREDICT_EXACT_DIV=1", "-c", os.path.join(_build.CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([_build._nvcc(), "-shared", "-cudart", "static", "-o", AB] + objs, check=True)
    print(AB)


def worker(out):
    import copy
    import bench
    import kadal_b200 as kb
    from oracle import kriging_oracle as ko
    res = {}
    rng = np.random.default_rng(7)
    X4 = rng.uniform(-1, 1, (800, 20)); y4 = (np.sin(X4[:, :3].sum(1)) + 0.05 * rng.standard_normal(800))[:, None] + 3.0
    X3, y3, th3 = bench.c3_data()
    for name, X, y, th in (("c3", X3, y3, th3), ("c4", X4, y4, np.full(20, 0.4))):
        info = ko.make_kriginfo(X, y)
        ref = copy.deepcopy(info)
        kb.likelihood(th.copy(), info, "all", False)
        ko.likelihood(th.copy(), ref, "all", False, check_pd=False)
        xp = np.random.default_rng(3).uniform(-1, 1, (4000, X.shape[1]))
        f, ss = kb.prediction(xp, info, ["pred", "SSqr"])
        fr, sr = ko.prediction(xp, ref, ["pred", "SSqr"])
        s2 = float(ref["SigmaSqr"][0, 0])
        res[name] = {"mean_rel": float(np.max(np.abs(f - fr) / np.maximum(np.abs(fr), 1))),
                     "var_rel": float(np.max(np.abs(ss - sr) / np.maximum(np.abs(sr), 1e-6 * s2))),
                     "mean_rel_median": float(np.median(np.abs(f - fr) / np.maximum(np.abs(fr), 1))),
                     "cond": float(np.linalg.cond(ref["Psi"] + 1e-6 * np.eye(len(X))))}
        np.save(out + "_" + name + ".npy", np.c_[f, ss])
    json.dump(res, open(out + ".json", "w"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    elif len(sys.argv) > 2 and sys.argv[1] == "worker":
        worker(sys.argv[2])
    else:
        tmp = os.path.join(ROOT, "gpurun_out")
        os.makedirs(tmp, exist_ok=True)
        for tag, env in (("recip", {}), ("exact", {"KB200_LIB": AB})):
            subprocess.run([sys.executable, os.path.abspath(__file__), "worker", os.path.join(tmp, "ab_" + tag)], check=True, env=dict(os.environ, **env))
        a, b = (json.load(open(os.path.join(tmp, "ab_%s.json" % t))) for t in ("recip", "exact"))
        for name in ("c3", "c4"):
            fa, fb = (np.load(os.path.join(tmp, "ab_%s_%s.npy" % (t, name))) for t in ("recip", "exact"))
            d = np.abs(fa - fb) / np.maximum(np.abs(fb), [1.0, 1e-300])
            print(name, "cond %.1e | vs oracle: reciprocal mean %.2e var %.2e | exact division mean %.2e var %.2e | between the two builds: mean %.2e var %.2e" % (
                a[name]["cond"], a[name]["mean_rel"], a[name]["var_rel"], b[name]["mean_rel"], b[name]["var_rel"], d[:, 0].max(), np.median(d[:, 1])))
