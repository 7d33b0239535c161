"""ms per 512-candidate C2 generation for every correlation kernel / composite (assembly path coverage)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kadal_b200.model import GpuModel
X, y, pop = bench.c2_data()
n, P = 1000, 512
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
rng = np.random.default_rng(1)
for name, kids, extra in (("gaussian", [0], 0), ("exponential", [1], 0), ("matern32", [2], 0), ("matern52", [3], 0),
                          ("gaussian+matern32", [0, 2], 2), ("gaussian d=10 trainvar", [0], -1)):
    m = GpuModel(X, y, np.ones((n, 1)), kids)
    pp = pop if extra <= 0 else np.hstack([pop, rng.uniform(0.2, 1, (P, extra))])
    tv = extra == -1
    if tv: pp = np.hstack([pop, rng.uniform(-1, 1, (P, 1))])
    xd = torch.from_numpy(np.ascontiguousarray(pp)).to(dev); nll = torch.empty(P, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
    f = lambda: m.lik_batch_dev(xd.data_ptr(), P, pp.shape[1], tv, -6.0, nll.data_ptr(), info.data_ptr(), st.cuda_stream)
    for _ in range(2): f()
    torch.cuda.synchronize()
    m.profile(True) if hasattr(m, "profile") else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    print("%-24s %.2f ms/generation  bad %d" % (name, e0.elapsed_time(e1) / 3, int((info != 0).sum().item())), flush=True)
    m.close()
