"""One full device refit and one one-sample append at the same hyper-parameters (for an ncu launch list).
    python tools/prof_append.py [n] [d]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from kadal_b200.model import GpuModel   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (n + 1, d))
y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n + 1) + 3.0
x = np.full(d, 0.3)
prev = GpuModel(X[:n], y[:n], np.ones((n, 1)), [0])
prev.fit_state(x, False, -6.0, want_U=False, want_Psi=False)
nxt = GpuModel(X, y, np.ones((n + 1, 1)), [0])
a = nxt.fit_state(x, False, -6.0, want_U=False, want_Psi=False, append_to=prev)
b = nxt.fit_state(x, False, -6.0, want_U=False, want_Psi=False)
print(a["nll"], b["nll"], a["nll"] - b["nll"])
