"""C2 shape (n=1000, d=10), candidates per GPU of a strong-scaled 512-candidate population (512, 256, 128, 64, 32):
ms per batch for the left-looking and the right-looking schedule, and for 1..4 stream groups."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kadal_b200.model import GpuModel
X, y, popB = bench.c2_data()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
for P in (512, 256, 128, 64, 32):
    row = []
    for name, env in (("left g2", {"KB200_RL": "0", "KB200_GROUPS": "2"}), ("left g1", {"KB200_RL": "0", "KB200_GROUPS": "1"}),
                      ("left g4", {"KB200_RL": "0", "KB200_GROUPS": "4"}), ("right g2", {"KB200_RL": "1", "KB200_GROUPS": "2"}),
                      ("right g1", {"KB200_RL": "1", "KB200_GROUPS": "1"})):
        for k in ("KB200_RL", "KB200_GROUPS", "KB200_GROUP_MIN"): os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["KB200_GROUP_MIN"] = "8"
        m = GpuModel(X, y, np.ones((len(X), 1)), [0])
        pop = popB[:P]
        xd = torch.from_numpy(pop).to(dev); nll = torch.empty(P, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
        f = lambda: m.lik_batch_dev(xd.data_ptr(), P, 10, False, -6.0, nll.data_ptr(), info.data_ptr(), st.cuda_stream)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        row.append("%s %.3f" % (name, e0.elapsed_time(e1) / 10))
        m.close()
    print("P", P, "|", "  ".join(row), flush=True)
