"""Randomised self-consistency stress of the round-2 kernels (no oracle, sizes the oracle would take minutes for):
  * lik_small_kernel vs the tile path on random (n <= 192, d, P, kernel lists, trainvar): bit-equal for n <= 128, 1e-9 beyond;
    repeated calls bit-identical (no race between the assembly, the factorisation and the right-hand-side strip);
  * right-looking prediction sweep vs the left-looking one on random (n, npred): variance bit-equal, mean to 1e-13;
  * device Sobol generator vs scipy at random offsets."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kadal_b200 as kb
from kadal_b200.model import GpuModel
from kadal_b200 import _capi

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
KERNELS = [[0], [1], [2], [3], [0, 2], [0, 1, 3]]
bad = 0
for it in range(60):
    n = int(rng.integers(1, 193)); d = int(rng.integers(1, 11)); P = int(rng.integers(1, 700))
    kid = KERNELS[int(rng.integers(0, len(KERNELS)))]
    tv = bool(rng.integers(0, 2)) and len(kid) == 1
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1)) + 3.0
    m = GpuModel(X, y, np.ones((n, 1)), kid)
    pop = rng.uniform(-0.7, 1.0, (P, d))
    if tv: pop = np.hstack([pop, rng.uniform(-1, 1, (P, 1))])
    if len(kid) > 1: pop = np.hstack([pop, rng.uniform(0.1, 1, (P, len(kid)))])
    a, ia = m.lik_batch(pop, tv, -6.0)
    a2, _ = m.lik_batch(pop, tv, -6.0)
    os.environ["KB200_NO_SMALL_LIK"] = "1"
    b, ib = m.lik_batch(pop, tv, -6.0)
    del os.environ["KB200_NO_SMALL_LIK"]
    ok = np.array_equal(a, a2) and np.array_equal(ia, ib)
    fin = np.isfinite(a) & np.isfinite(b)
    if n <= 128: ok = ok and np.array_equal(a[fin], b[fin]) and np.array_equal(np.isfinite(a), np.isfinite(b))
    else: ok = ok and np.all(np.abs(a[fin] - b[fin]) <= 2e-9 * np.maximum(1, np.abs(b[fin]))) and np.array_equal(np.isfinite(a), np.isfinite(b))
    if not ok:
        bad += 1
        print("LIK MISMATCH n", n, "d", d, "P", P, "kid", kid, "tv", tv, np.max(np.abs(a[fin] - b[fin])) if fin.any() else None, flush=True)
    m.close()
print("lik_small stress done, mismatches:", bad, flush=True)
for it in range(12):
    n = int(rng.integers(257, 1400)); d = int(rng.integers(1, 12)); npred = int(rng.integers(1, 6000))
    X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1)) + 3.0
    m = GpuModel(X, y, np.ones((n, 1)), [0])
    st = m.fit_state(np.full(d, 0.1), False, -6.0, want_U=False, want_Psi=False)
    assert st["info"] == 0
    xp = rng.uniform(-1, 1, (npred, d))
    lb, ub = -np.ones(d), np.ones(d)
    res = {}
    for mode in ("1", "0"):
        os.environ["KB200_PRED_RL"] = mode
        res[mode] = m.predict(xp, _capi.OUT_PRED | _capi.OUT_SSQR, _capi.NORM_DEFAULT, lb, ub)
    del os.environ["KB200_PRED_RL"]
    ok = np.array_equal(res["1"]["ssqr"], res["0"]["ssqr"]) and np.all(np.abs(res["1"]["pred"] - res["0"]["pred"]) <= 1e-12 * np.maximum(1, np.abs(res["0"]["pred"])))
    if not ok:
        bad += 1
        print("PRED MISMATCH n", n, "d", d, "npred", npred, np.max(np.abs(res["1"]["ssqr"] - res["0"]["ssqr"])), flush=True)
    m.close()
print("prediction sweep stress done, mismatches so far:", bad, flush=True)
import torch
from kadal_b200 import sampling as sp
from scipy.stats import qmc
for it in range(6):
    dims = int(rng.integers(1, 64)); i0 = int(rng.integers(0, 2 ** 20)); n = int(rng.integers(1, 50000))
    out = torch.empty((n, dims), dtype=torch.float64, device="cuda:0")
    if not sp.sobol_points_device(n, dims, first_row=i0, device=0, out_ptr=out.data_ptr()):
        break
    eng = qmc.Sobol(dims, scramble=False); eng.fast_forward(i0)
    if not np.array_equal(out.cpu().numpy(), eng.random(n)):
        bad += 1
        print("SOBOL MISMATCH", dims, i0, n, flush=True)
print("stress ok" if bad == 0 else "stress FAILED: %d" % bad)
sys.exit(1 if bad else 0)
