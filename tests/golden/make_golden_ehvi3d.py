"""Generate tests/golden/ehvi3d.npz from the reference's own KMAC C++ (oracle/_ref/libkmac_ref.so, compiled from
/root/reference/kadal/extern/HDYE_3D_Update by oracle/kmac_ref/Makefile; build container only):

    PYTHONPATH=/root/repo python tests/golden/make_golden_ehvi3d.py

Cases: the reference's two known-answer tests (test_single.py, test_multi.py: inputs typed in here as data,
answers recomputed and checked against the asserted values), then random mutually non-dominated fronts with
mixed candidates.  `v` = the batch as the reference scores it (order-dependent), `w` = every candidate scored
alone."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import kmac_ref_loader as kr  # noqa: E402

KAT_SINGLE_FRONT = np.array([
    [-0.981894, -9.951677, -4.021090], [-1.675298, -9.858670, -3.970364], [-6.989333, -7.151868, -6.273483],
    [-6.524929, -7.577948, -6.841600], [-9.967894, -0.800677, -4.963238], [-7.503389, -6.610533, -8.704238],
    [-8.988774, -4.382002, -4.607492], [-9.909816, -1.339981, -6.117480], [-8.487371, -5.288151, -6.438087],
    [-4.616053, -8.870854, -1.834406]])
KAT_SINGLE_MEAN = np.array([[-7.255624, -6.881563, -4.043903]])
KAT_SINGLE_ANSWER = 9.739142427674365                     # test_single.py:21
KAT_MULTI_FRONT = np.array([
    [-4.632504, -8.862274, -5.369237], [-1.748896, -9.845880, -3.447994], [-7.108236, -7.033704, -9.625512],
    [-6.209090, -7.838827, -6.780259], [-6.446806, -7.644520, -2.314544], [-5.243587, -8.514975, -1.138873],
    [-2.310623, -9.729389, -9.705481], [-9.573677, -2.888719, -0.198476], [-9.996871, -0.250155, -4.509609],
    [-1.022884, -9.947548, -4.678429]])
KAT_MULTI_MEAN = np.array([[-2.917795, -9.564856, -3.172051], [-2.9, -9.5, -3.2], [-6.605903, -7.507466, -1.988600]])
KAT_MULTI_ANSWER = np.array([3.707616667, 4.563448328, 0])  # test_multi.py:26
KAT_REF = np.array([-10.0, -10.0, -10.0])


def random_front(rng, k):
    """k points on the positive octant of a sphere: mutually non-dominated under maximisation."""
    v = np.abs(rng.standard_normal((k, 3)))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return -10 + 9.5 * v


def cases():
    rng = np.random.default_rng(20261020)
    out = [(KAT_SINGLE_FRONT, KAT_REF, KAT_SINGLE_MEAN, np.zeros((1, 3))),
           (KAT_MULTI_FRONT, KAT_REF, KAT_MULTI_MEAN, np.zeros((3, 3)))]
    for k, npop, zero in ((0, 9, True), (1, 17, False), (2, 17, True), (5, 40, False), (10, 64, True), (17, 64, False),
                          (40, 96, False), (40, 96, True)):
        P = random_front(rng, k) if k else np.zeros((0, 3))
        mu = rng.uniform(-10.5, -0.5, (npop, 3))
        s = rng.uniform(0.01, 2.5, (npop, 3))
        if k:
            mu[3] = P[0]                 # a candidate sitting on a front point
            P = np.vstack([P, P[0] - 0.3])   # a dominated point in the input: later points do not remove earlier ones
        s[npop // 2:, :] *= 0.05         # sharp candidates: many negligible cells -> the early exit matters
        if zero:
            s[5] = 0.0                   # deterministic candidate (every cell but one is negligible for it)
            s[7, 1] = 0.0
        out.append((P, np.array([-10.0, -10.2, -10.4]), mu, s))
    return out


if __name__ == "__main__":
    assert kr.build() and kr.available()
    res = {}
    cs = cases()
    for c, (P, r, mu, s) in enumerate(cs):
        v = kr.ehvi3d_multi(P, r, mu, s)
        w = np.array([kr.ehvi3d_multi(P, r, mu[i:i + 1], s[i:i + 1])[0] for i in range(len(mu))])
        res.update({"P%d" % c: P, "r%d" % c: r, "mu%d" % c: mu, "s%d" % c: s, "v%d" % c: v, "w%d" % c: w})
        print(c, "k", len(P), "npop", len(mu), "batch max %.4g  alone max %.4g  candidates that differ: %d" % (
            v.max(), w.max(), int((v != w).sum())))
    assert np.isclose(res["v0"][0], KAT_SINGLE_ANSWER) and np.isclose(res["v1"], KAT_MULTI_ANSWER).all()
    res["ncases"] = len(cs)
    np.savez_compressed(os.path.join(HERE, "ehvi3d.npz"), **res)
