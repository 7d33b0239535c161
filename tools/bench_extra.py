"""Throughput of the other BASELINE configs (device-resident inputs, CUDA events on the library's
stream), one JSON object per line:

    python tools/bench_extra.py [c4 c4var c5lik c5pred c3n300]

  c4      C4 Sobol: n=800, d=20, 'pred' only over 2e6 points        (corr_tile_kernel, predict mode)
  c4var   n=800, d=20, pred + s over 4e5 points                      (general tiled variance path)
  c3n300  n=300, d=6, pred + s over 2e6 points                       (n > 256: general path)
  c5lik   C5 KPLS likelihood: n=4000, d=100, h=3, 16 candidates      (blocked tensor-core Cholesky)
  c5pred  C5 candidate scoring: pred + SSqr over 1e5 points
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kadal_b200 import _capi  # noqa: E402
from kadal_b200.model import GpuModel, fp64_probe  # noqa: E402


def timed(torch, fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pred_case(torch, name, X, y, th, npts, want, pls=None, reps=3, kernels=(0,)):
    dev = torch.device("cuda", 0)
    n, d = X.shape
    m = GpuModel(X, y, np.ones((n, 1)), list(kernels), pls=pls)
    st = m.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
    assert st["info"] == 0, st["info"]
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    xs = (torch.rand(npts, d, generator=g, device=dev, dtype=torch.float64) * 2 - 1)
    outs = [torch.empty(npts, dtype=torch.float64, device=dev) for _ in range(7)]
    stream = torch.cuda.current_stream().cuda_stream
    ptrs = [outs[i].data_ptr() if want & (1 << i) else 0 for i in range(7)]
    lb, ub = -np.ones(d), np.ones(d)

    def step():
        m.predict_dev(xs.data_ptr(), npts, want, ptrs, normtype=_capi.NORM_DEFAULT, norm_a=lb, norm_b=ub, stream=stream)

    ms = timed(torch, step, reps)
    m.profile(True)
    step()
    torch.cuda.synchronize()
    m.profile(False)
    prof = {k: round(v[0], 3) for k, v in m.profile_get().items() if v[1]}
    var = bool(want & ~_capi.OUT_PRED)
    nth = pls.shape[1] if pls is not None else d
    fl = bench.predict_flops_per_point(n, d, var)
    pts = npts / (ms * 1e-3)
    m.close()
    return {"case": name, "n": n, "d": d, "npts": npts, "ms": ms, "pts_per_s": pts, "kernel_ms": prof,
            "fp64_tflops_algorithmic": pts * fl / 1e12, "ntheta": nth}


def main():
    import torch
    cases = sys.argv[1:] or ["c4", "c4var", "c3n300", "c5lik", "c5pred"]
    tstream = torch.cuda.Stream(torch.device("cuda", 0))
    torch.cuda.set_stream(tstream)
    dmma, dfma = fp64_probe(0)
    print(json.dumps({"dmma_tflops": dmma, "dfma_tflops": dfma}), flush=True)
    for c in cases:
        if c in ("c4", "c4var"):
            rng = np.random.default_rng(7)
            X = rng.uniform(-1, 1, (800, 20))
            y = (np.sin(X[:, :3].sum(1)) + 0.05 * rng.standard_normal(800))[:, None] + 3.0
            th = np.full(20, 0.4)
            if c == "c4":
                r = pred_case(torch, c, X, y, th, 2_000_000, _capi.OUT_PRED)
            else:
                r = pred_case(torch, c, X, y, th, 400_000, _capi.OUT_PRED | _capi.OUT_S)
        elif c == "c3":      # C3 shape; $KB200_NO_FUSED_PREDICT=1 sends it through the general tiled path instead of predict_small_kernel
            X, y, th = bench.c3_data()
            r = pred_case(torch, c, X, y, th, 4_000_000, _capi.OUT_PRED | _capi.OUT_S)
        elif c in ("c3m52", "c3d10"):   # small models outside the fused predictor's envelope: Matern 5/2, or d = 10 -> general tiled path
            if c == "c3m52":
                X, y, th = bench.c3_data()
                r = pred_case(torch, c, X, y, th, 4_000_000, _capi.OUT_PRED | _capi.OUT_S, kernels=(3,))
            else:
                rng = np.random.default_rng(1)
                X = rng.uniform(-1, 1, (200, 10))
                y = (np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2)[:, None] + 3.0
                r = pred_case(torch, c, X, y, np.full(10, 0.3), 4_000_000, _capi.OUT_PRED | _capi.OUT_S)
        elif c == "c3n300":
            rng = np.random.default_rng(1)
            X = rng.uniform(-1, 1, (300, 6))
            y = (np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2)[:, None] + 3.0
            r = pred_case(torch, c, X, y, np.full(6, 0.2), 2_000_000, _capi.OUT_PRED | _capi.OUT_S)
        elif c in ("c5lik", "c5pred"):
            rng = np.random.default_rng(11)
            n, d, h = 4000, 100, 3
            X = rng.uniform(-1, 1, (n, d))
            y = (X[:, :5].sum(1) ** 2)[:, None] + 3.0
            pls = rng.standard_normal((d, h)) / 10
            th = np.array([0.3, 0.1, -0.1])
            if c == "c5pred":
                r = pred_case(torch, c, X, y, th, 100_000, _capi.OUT_PRED | _capi.OUT_SSQR, pls=pls, reps=2)
            else:
                P = 16
                m = GpuModel(X, y, np.ones((n, 1)), [0], pls=pls)
                pop = th[None, :] + np.random.default_rng(3).uniform(-0.2, 0.2, (P, h))
                dev = torch.device("cuda", 0)
                xd = torch.from_numpy(pop).to(dev)
                nll = torch.empty(P, dtype=torch.float64, device=dev)
                info = torch.empty(P, dtype=torch.int32, device=dev)
                stream = torch.cuda.current_stream().cuda_stream

                def step():
                    m.lik_batch_dev(xd.data_ptr(), P, h, False, -6.0, nll.data_ptr(), info.data_ptr(), stream)

                ms = timed(torch, step, 2)
                m.set_groups(1)
                m.profile(True)
                step()
                torch.cuda.synchronize()
                m.profile(False)
                prof = {k: round(v[0], 3) for k, v in m.profile_get().items() if v[1]}
                assert int((info != 0).sum().item()) == 0
                fl = n ** 3 / 3.0 * P
                r = {"case": c, "n": n, "d": d, "P": P, "ms": ms, "evals_per_s": P / (ms * 1e-3), "kernel_ms": prof,
                     "chol_tflops": fl / (ms * 1e-3) / 1e12}
                m.close()
        else:
            continue
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
