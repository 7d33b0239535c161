"""ms per 512-candidate C2 generation (gaussian) and the per-kernel-class times of a serialised pass."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kadal_b200.model import GpuModel
X, y, pop = bench.c2_data()
n, P = 1000, 512
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
m = GpuModel(X, y, np.ones((n, 1)), [0])
xd = torch.from_numpy(np.ascontiguousarray(pop)).to(dev); nll = torch.empty(P, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
f = lambda: m.lik_batch_dev(xd.data_ptr(), P, pop.shape[1], False, -6.0, nll.data_ptr(), info.data_ptr(), st.cuda_stream)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / 10
ref = nll.cpu().numpy().copy()
m.set_groups(1); m.profile(True)
for _ in range(5): f()
torch.cuda.synchronize()
pr = m.profile_get()
print("%s  %.3f ms/generation  assemble %.3f ms  diag %.3f  panel %.3f  bad %d  nll_sum %.12e" % (
    os.environ.get("KB200_NVCC_EXTRA", "default"), tot, pr["assemble"][0] / 5, pr["chol_diag"][0] / 5, pr["chol_panel"][0] / 5,
    int((info != 0).sum().item()), float(ref.sum())), flush=True)

# mean-only prediction through corr_tile_kernel (predict mode): d = 6 (n = 300) and d = 10 (the C2 model)
m.set_groups(2); m.profile(False)
for (nn, dd) in (() if os.environ.get("KB200_TIME_C2_ONLY") else ((300, 6), (1000, 10))):
    rng = np.random.default_rng(3)
    Xs = rng.uniform(-1, 1, (nn, dd)); ys = np.sin(Xs.sum(1)) + 3.0
    mm = GpuModel(Xs, ys, np.ones((nn, 1)), [0])
    mm.fit_state(np.full(dd, 0.2), False, -6.0, want_U=False, want_Psi=False)
    npts = 2_000_000
    xq = torch.from_numpy(rng.uniform(-1, 1, (npts, dd))).to(dev)
    out = torch.empty(npts, dtype=torch.float64, device=dev)
    g = lambda: mm.predict_dev(xq.data_ptr(), npts, 1, [out.data_ptr(), 0, 0, 0, 0, 0, 0], normtype=1, norm_a=-np.ones(dd), norm_b=np.ones(dd), stream=st.cuda_stream)
    for _ in range(2): g()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5): g()
    e1.record(); torch.cuda.synchronize()
    print("    pred n=%d d=%d: %.3f ms per 2e6 points, sum %.12e" % (nn, dd, e0.elapsed_time(e1) / 5, float(out.sum().item())), flush=True)
    mm.close()
