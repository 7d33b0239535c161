import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kadal_b200 import ehvi
rng = np.random.default_rng(5)
k, npop = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
v = np.abs(rng.standard_normal((k, 3))); P = -10 + 9.5 * v / np.linalg.norm(v, axis=1, keepdims=True)
mu, sd = rng.uniform(-10.2, -0.5, (npop, 3)), rng.uniform(0.3, 2.5, (npop, 3))
for mode in ("reference", "independent"):
    for _ in range(2):
        h = ehvi.ehvi3d_sliceupdate_multi(P, [-10.0] * 3, mu, sd, mode=mode)
    print(mode, int((h > 0).sum()), float(h.max()))
