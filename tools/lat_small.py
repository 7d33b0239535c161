import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from kadal_b200.model import GpuModel
X, y, pop = bench.c2_data()
for n in (200, 1000):
    for h in ("0", "1"):
        os.environ["KB200_HALF"] = h
        m = GpuModel(X[:n], y[:n], np.ones((n, 1)), [0])
        for P in (1, 8, 64):
            for _ in range(3): m.lik_batch(pop[:P], False, -6.0)
            t0 = time.perf_counter()
            for _ in range(20): m.lik_batch(pop[:P], False, -6.0)
            print("n", n, "half", h, "P", P, "ms/call %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
        m.close()
