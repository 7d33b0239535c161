"""Binary128 (IEEE quad, libquadmath) value of the reference's ln-likelihood formula for EVERY candidate of the two
C2 populations (2 x 512), from the SAME float64 Psi the reference builds (oracle/truth/nll_quad.c):

    make -C oracle/truth && OMP_NUM_THREADS=1 PYTHONPATH=/root/repo python tests/golden/make_truth_full.py

Writes tests/golden/c2_truth_full.npz: truth_popA, truth_popB (float64 roundings of the binary128 results) and ref1t_*
(the float64 formula evaluated here with ONE BLAS thread: it differs from the two-thread run of c2_full.npz by up to
2.1e-8 relative on popB -- the reference does not even agree with itself to 1e-9 across thread counts).  The float64
results of the live reference for the same candidates are in c2_full.npz (nll_popA / nll_popB, cond_*).  Together they
let the GPU parity test assert, candidate by candidate: within 1e-9 of the reference, OR closer to the exact value of the
formula than the reference itself is ("what's weak" 1).  ~3-6 s per candidate and core."""
import ctypes, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import kriging_oracle as ko  # noqa: E402


def _lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtruth_quad.so"))
    lib.kb_truth_nll_quad.restype = ctypes.c_double
    lib.kb_truth_nll_quad.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib


def one(job):
    pop, i = job
    w = ko.workload("C2")
    info = ko.make_kriginfo(w["X"], w["y"])
    x = w[pop][i].copy()
    ko.likelihood(x, info, "all", False, check_pd=False)          # fills info['Psi'] (without nugget), float64 as the reference
    n = info["Psi"].shape[0]
    A = np.ascontiguousarray(info["Psi"] + np.eye(n) * 10.0 ** info["nugget"])   # likelihood_func.py:171 in float64
    y = np.ascontiguousarray(w["y"].ravel())
    return pop, i, _lib().kb_truth_nll_quad(A.ctypes.data, y.ctypes.data, n), float(info["NegLnLike"])


if __name__ == "__main__":
    from multiprocessing import get_context
    jobs = [(p, i) for p in ("popA", "popB") for i in range(512)]
    if len(sys.argv) > 1:
        jobs = jobs[:int(sys.argv[1])]
    with get_context("spawn").Pool(int(os.environ.get("TRUTH_PROCS", "7"))) as pool:
        res = pool.map(one, jobs, chunksize=4)
    out = {"truth_popA": np.full(512, np.nan), "truth_popB": np.full(512, np.nan),
           "ref1t_popA": np.full(512, np.nan), "ref1t_popB": np.full(512, np.nan)}
    for pop, i, t, r in res:
        out["truth_" + pop][i] = t
        out["ref1t_" + pop][i] = r      # the float64 formula again, single-threaded BLAS (c2_full.npz: live reference, 2 threads)
    np.savez_compressed(os.path.join(HERE, "c2_truth_full.npz"), **out)
    g = np.load(os.path.join(HERE, "c2_full.npz"))
    for pop in ("popA", "popB"):
        ok = np.isfinite(out["truth_" + pop])
        d = np.abs(g["nll_" + pop] - out["truth_" + pop])[ok] / np.maximum(1, np.abs(out["truth_" + pop][ok]))
        d1 = np.abs(out["ref1t_" + pop] - g["nll_" + pop])[ok] / np.maximum(1, np.abs(g["nll_" + pop][ok]))
        print(pop, "n", ok.sum(), "max |ref-truth| rel %.2e" % d.max(), "in cond*ulp: max %.3f" % (d / (g["cond_" + pop][ok] * 2.0 ** -53)).max(),
              "reference, 1 BLAS thread vs 2: max rel %.2e" % d1.max())
