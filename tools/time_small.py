"""ms per 512-candidate likelihood batch at small n: lik_small_kernel (one CTA per candidate in shared memory) against the
tile path ($KB200_NO_SMALL_LIK=1) the same sizes took before; algorithmic TFLOP/s beside it."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kadal_b200.model import GpuModel
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for n, d in ((50, 2), (64, 4), (100, 6), (128, 6), (160, 6), (192, 6), (200, 6), (224, 6), (200, 10)):
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1))[:, None] + 3.0
    pop = rng.uniform(-0.5, 1.0, (P, d))
    m = GpuModel(X, y, np.ones((n, 1)), [0])
    xd = torch.from_numpy(pop).to(dev); nll = torch.empty(P, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
    f = lambda: m.lik_batch_dev(xd.data_ptr(), P, d, False, -6.0, nll.data_ptr(), info.data_ptr(), st.cuda_stream)
    row = []
    for name, env in (("small", None), ("tile", "1")):
        if env: os.environ["KB200_NO_SMALL_LIK"] = env
        else: os.environ.pop("KB200_NO_SMALL_LIK", None)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        row.append("%s %.4f ms (%.2f TF)" % (name, ms, bench.lik_flops_per_eval(n, d) * P / (ms * 1e-3) / 1e12))
    os.environ.pop("KB200_NO_SMALL_LIK", None)
    print("n %3d d %2d P %d | %s" % (n, d, P, "   ".join(row)), flush=True)
    m.close()
