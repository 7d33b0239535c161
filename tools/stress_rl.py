"""Determinism stress of the right-looking schedule (look-ahead on two streams): repeated small batches must
give the same bits every time and agree with the left-looking schedule to rounding."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from kadal_b200.model import GpuModel
X, y, pop = bench.c2_data()
n = 1000
os.environ["KB200_RL"] = "0"
m0 = GpuModel(X, y, np.ones((n, 1)), [0]); ref, _ = m0.lik_batch(pop[:32], False, -6.0); m0.close()
os.environ["KB200_RL"] = "1"
m = GpuModel(X, y, np.ones((n, 1)), [0])
bad = 0
base = {}
for rep in range(150):
    P = (1, 3, 16, 32)[rep % 4]
    nll, inf = m.lik_batch(pop[:P], False, -6.0)
    if P not in base: base[P] = nll.copy()
    bad += int((nll != base[P]).sum()) + int(inf.any())
# sub-batches of the right-looking schedule are bit-identical too
bad += int((base[32][:16] != base[16]).sum()) + int(base[1][0] != base[32][0])
print("right-looking: bad or non-reproducible", bad, "max rel vs left-looking", float(np.max(np.abs(base[32] - ref) / np.maximum(1, np.abs(ref)))))
