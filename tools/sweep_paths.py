"""ms per likelihood batch for the three factorisation paths at several (n, P): validates the size policy."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from kadal_b200.model import GpuModel
rng = np.random.default_rng(5)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
for n, P in ((1000, 16), (1000, 64), (1537, 16), (1537, 64), (2048, 8), (2048, 32), (2048, 128), (3000, 16), (4000, 64)):
    d = 8
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1))[:, None] + 0.1 * rng.standard_normal((n, 1))
    pop = rng.uniform(0.0, 1.0, (P, d))
    row = []
    for name, env in (("full", {"KB200_HALF": "0"}), ("half", {"KB200_HALF": "1", "KB200_RL": "0"}), ("rl", {"KB200_HALF": "1", "KB200_RL": "1"})):
        for k in ("KB200_HALF", "KB200_RL"): os.environ.pop(k, None)
        os.environ.update(env)
        m = GpuModel(X, y, np.ones((n, 1)), [0])
        xd = torch.from_numpy(pop).to(dev); nll = torch.empty(P, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
        f = lambda: m.lik_batch_dev(xd.data_ptr(), P, d, False, -6.0, nll.data_ptr(), info.data_ptr(), st.cuda_stream)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        row.append("%s %.2f" % (name, e0.elapsed_time(e1) / 3))
        m.close()
    print("n", n, "P", P, "|", "  ".join(row), flush=True)
