"""Stress test of the half-SM factorisation path: repeated 148..512-candidate C2 batches must be bit-identical
from run to run and agree with the full-SM path (DBG_P, DBG_REPS, DBG_N; see profiles/r1k_half_race.md)."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from kadal_b200.model import GpuModel
X, y, pop = bench.c2_data()
n = int(os.environ.get("DBG_N", "1000")); P = int(os.environ.get("DBG_P", "512")); reps = int(os.environ.get("DBG_REPS", "3"))
X, y = X[:n], y[:n]
os.environ["KB200_HALF"] = "0"
m0 = GpuModel(X, y, np.ones((n, 1)), [0]); n0, i0 = m0.lik_batch(pop[:P], False, -6.0); m0.close()
os.environ["KB200_HALF"] = "1"
m = GpuModel(X, y, np.ones((n, 1)), [0])
tot = 0
first = None
for rep in range(reps):
    n1, i1 = m.lik_batch(pop[:P], False, -6.0)
    if first is None: first = n1.copy()
    tot += int((i1 != 0).sum()) + int((n1 != first).sum())
ok = (i1 == 0) & (i0 == 0)
print("reps", reps, "P", P, "total bad or non-reproducible", tot, "maxrel vs full-tile path", float(np.max(np.abs(n1[ok]-n0[ok])/np.maximum(1,np.abs(n0[ok])))), flush=True)
