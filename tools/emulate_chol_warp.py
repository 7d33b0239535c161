"""Lane-level numpy emulation of the warp algorithms in kb200_kernels.cu (development aid).

Checks the index arithmetic of the blocked in-tile POTF2 (potf2_32_warp, inv8_warp,
strip_trsm32, trailing_update32) and of the panel kernel's right-looking TRSM against
numpy.linalg, using the documented mma.sync.m8n8k4.f64 fragment layout
    a = A[g][t]   b = B[t][g]   (c0, c1) = C[g][2t], C[g][2t+1]     g = lane/4, t = lane%4.
Run:  python tools/emulate_chol_warp.py
"""
import numpy as np

TB, LDP, TLD, KS = 128, 129, 132, 16
LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def dmma(c0, c1, a, b):
    """c0,c1,a,b: arrays over 32 lanes. Returns updated (c0,c1)."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[G, T] = a
    B[T, G] = b
    C[G, 2 * T] = c0
    C[G, 2 * T + 1] = c1
    D = A @ B + C
    return D[G, 2 * T], D[G, 2 * T + 1]


def shfl(v, src):
    return v[src]


def cfrag_to_afrag(c0, c1):
    src0 = (LANES & ~3) + ((LANES & 3) >> 1)
    v00, v01 = shfl(c0, src0), shfl(c1, src0)
    v10, v11 = shfl(c0, src0 + 2), shfl(c1, src0 + 2)
    a0 = np.where(LANES & 1, v01, v00)
    a1 = np.where(LANES & 1, v11, v10)
    return a0, a1


def potf2_32_warp(S, o, dinv):
    a = np.zeros((32, 32))       # a[lane][c]
    for c in range(32):
        a[:, c] = np.where(c <= LANES, S[o + LANES, o + c], 0.0)
    for k in range(32):
        piv = a[k, k]
        inv = 1.0 / np.sqrt(piv)
        dk = piv * inv
        l = a[:, k] * inv
        for c in range(k + 1, 32):
            lc = l[c]
            a[:, c] = a[:, c] - l * lc
        a[:, k] = np.where(LANES == k, dk, l)
        dinv[o + k] = inv
    for c in range(32):
        m = c <= LANES
        S[o + LANES[m], o + c] = a[m, c]


def inv8_warp(S, o, dinv, W8, kb):
    blk, c = LANES >> 3, LANES & 7
    w = np.zeros((32, 8))
    for r in range(8):
        s = np.zeros(32)
        for t in range(r):
            s = s + S[o + 8 * blk + r, o + 8 * blk + t] * w[:, t]
        w[:, r] = np.where(r == c, 1.0, -s) * dinv[o + 8 * blk + r]
        w[:, r] = np.where(r < c, 0.0, w[:, r])
    for r in range(8):
        W8[kb * 256 + blk * 64 + r * 8 + c] = w[:, r]


def strip_trsm32(rows, r0, c0, Sfull, o, W8, kb):
    """rows: 2-D array holding the strip (rows r0..r0+7, columns c0..c0+31)."""
    x = [[rows[r0 + G, c0 + 8 * ct + 2 * T].copy(), rows[r0 + G, c0 + 8 * ct + 2 * T + 1].copy()] for ct in range(4)]
    for t in range(4):
        a0, a1 = cfrag_to_afrag(x[t][0], x[t][1])
        y0, y1 = np.zeros(32), np.zeros(32)
        y0, y1 = dmma(y0, y1, a0, W8[kb * 256 + t * 64 + G * 8 + T])
        y0, y1 = dmma(y0, y1, a1, W8[kb * 256 + t * 64 + G * 8 + 4 + T])
        x[t] = [y0, y1]
        if t < 3:
            a0, a1 = cfrag_to_afrag(y0, y1)
            a0, a1 = -a0, -a1
            for ct in range(t + 1, 4):
                b0 = Sfull[o + 8 * ct + G, o + 8 * t + T]
                b1 = Sfull[o + 8 * ct + G, o + 8 * t + 4 + T]
                x[ct][0], x[ct][1] = dmma(x[ct][0], x[ct][1], a0, b0)
                x[ct][0], x[ct][1] = dmma(x[ct][0], x[ct][1], a1, b1)
    for ct in range(4):
        rows[r0 + G, c0 + 8 * ct + 2 * T] = x[ct][0]
        rows[r0 + G, c0 + 8 * ct + 2 * T + 1] = x[ct][1]


def trailing_update32(S, trow, o, nwarp=16):
    base = o + 32
    nb = (TB - base) >> 4
    ntri = nb * (nb + 1) // 2
    for it in range(ntri + nb):
        if it < ntri:
            bi = 0
            while (bi + 1) * (bi + 2) // 2 <= it:
                bi += 1
            bj = it - bi * (bi + 1) // 2
            r0, r1 = base + 16 * bi + G, base + 16 * bi + 8 + G
            q0, q1 = base + 16 * bj + G, base + 16 * bj + 8 + G
            cc = base + 16 * bj + 2 * T
            c = [[[S[r0, cc + 8 * nt].copy(), S[r0, cc + 8 * nt + 1].copy()] for nt in range(2)],
                 [[S[r1, cc + 8 * nt].copy(), S[r1, cc + 8 * nt + 1].copy()] for nt in range(2)]]
            for kq in range(8):
                kc = o + 4 * kq + T
                a0, a1 = -S[r0, kc], -S[r1, kc]
                b0, b1 = S[q0, kc], S[q1, kc]
                c[0][0] = list(dmma(c[0][0][0], c[0][0][1], a0, b0))
                c[0][1] = list(dmma(c[0][1][0], c[0][1][1], a0, b1))
                c[1][0] = list(dmma(c[1][0][0], c[1][0][1], a1, b0))
                c[1][1] = list(dmma(c[1][1][0], c[1][1][1], a1, b1))
            for nt in range(2):
                if bi > bj or nt == 0:
                    S[r0, cc + 8 * nt] = c[0][nt][0]; S[r0, cc + 8 * nt + 1] = c[0][nt][1]
                S[r1, cc + 8 * nt] = c[1][nt][0]; S[r1, cc + 8 * nt + 1] = c[1][nt][1]
        else:
            bj = it - ntri
            q0, q1 = base + 16 * bj + G, base + 16 * bj + 8 + G
            cc = base + 16 * bj + 2 * T
            c = [[trow[G, cc + 8 * nt].copy(), trow[G, cc + 8 * nt + 1].copy()] for nt in range(2)]
            for kq in range(8):
                kc = o + 4 * kq + T
                a0 = -trow[G, kc]
                c[0] = list(dmma(c[0][0], c[0][1], a0, S[q0, kc]))
                c[1] = list(dmma(c[1][0], c[1][1], a0, S[q1, kc]))
            for nt in range(2):
                trow[G, cc + 8 * nt] = c[nt][0]; trow[G, cc + 8 * nt + 1] = c[nt][1]


def potf2_blocked(S, trow):
    dinv = np.zeros(TB)
    W8 = np.zeros(1024)
    for kb in range(4):
        o = 32 * kb
        potf2_32_warp(S, o, dinv)
        inv8_warp(S, o, dinv, W8, kb)
        nrow_strips = (TB - o - 32) >> 3
        for st in range(nrow_strips + 1):
            if st < nrow_strips:
                strip_trsm32(S, o + 32 + 8 * st, o, S, o, W8, kb)
            else:
                strip_trsm32(trow, 0, o, S, o, W8, kb)
        if kb < 3:
            trailing_update32(S, trow, o)
    return dinv, W8


def tile_off(r, c):
    return (c >> 4) * 2048 + r * 16 + ((((c >> 2) & 3) ^ (r & 3)) << 2) + (c & 3)


def panel_trsm(C, Ltile, W8):
    """C: 128x128 (row-major), Ltile: tile-layout L_jj (flat 16384), W8: 1024.  Returns X."""
    X = np.zeros((TB, TB))
    for warp in range(16):
        xr = 8 * warp + G
        x = [[C[xr, 8 * ct + 2 * T].copy(), C[xr, 8 * ct + 2 * T + 1].copy()] for ct in range(16)]
        for s in range(8):
            Ls = Ltile[s * 2048:(s + 1) * 2048]
            for tt in range(2):
                t = 2 * s + tt
                a0, a1 = cfrag_to_afrag(x[t][0], x[t][1])
                y0, y1 = np.zeros(32), np.zeros(32)
                y0, y1 = dmma(y0, y1, a0, W8[t * 64 + G * 8 + T])
                y0, y1 = dmma(y0, y1, a1, W8[t * 64 + G * 8 + 4 + T])
                x[t] = [y0, y1]
                if t < 15:
                    a0, a1 = cfrag_to_afrag(y0, y1)
                    a0, a1 = -a0, -a1
                    for ct in range(t + 1, 16):
                        lrow = (8 * ct + G) * KS + T
                        b0 = Ls[lrow + (((2 * tt) ^ (G & 3)) << 2)]
                        b1 = Ls[lrow + (((2 * tt + 1) ^ (G & 3)) << 2)]
                        x[ct][0], x[ct][1] = dmma(x[ct][0], x[ct][1], a0, b0)
                        x[ct][0], x[ct][1] = dmma(x[ct][0], x[ct][1], a1, b1)
        for ct in range(16):
            X[xr, 8 * ct + 2 * T] = x[ct][0]
            X[xr, 8 * ct + 2 * T + 1] = x[ct][1]
    return X


def main():
    rng = np.random.default_rng(0)
    M = rng.standard_normal((TB, TB))
    A = M @ M.T + TB * np.eye(TB)
    rhs = rng.standard_normal((8, TB))
    S = np.full((TB, LDP), np.nan)
    for r in range(TB):
        S[r, :r + 1] = A[r, :r + 1]
    trow = np.zeros((8, TLD))
    trow[:, :TB] = rhs
    dinv, W8 = potf2_blocked(S, trow)
    Lref = np.linalg.cholesky(A)
    L = np.tril(S[:, :TB])
    print("potf2 |L - Lref| max", np.abs(L - Lref).max())
    zref = np.linalg.solve(Lref, rhs.T).T
    print("rhs strip |z - zref| max", np.abs(trow[:, :TB] - zref).max())
    for b in range(16):
        blk = Lref[8 * b:8 * b + 8, 8 * b:8 * b + 8]
        assert np.allclose(W8[b * 64:(b + 1) * 64].reshape(8, 8), np.linalg.inv(blk), atol=1e-12), b
    print("W8 ok; dinv err", np.abs(dinv - 1 / np.diag(Lref)).max())
    # panel TRSM
    C = rng.standard_normal((TB, TB))
    Lt = np.zeros(TB * TB)
    for r in range(TB):
        for c in range(TB):
            Lt[tile_off(r, c)] = L[r, c]
    X = panel_trsm(C, Lt, W8)
    Xref = np.linalg.solve(Lref, C.T).T
    print("panel TRSM |X - Xref| max", np.abs(X - Xref).max())
    assert np.abs(L - Lref).max() < 1e-11 and np.abs(X - Xref).max() < 1e-11
    assert np.abs(trow[:, :TB] - zref).max() < 1e-11
    print("OK")


if __name__ == "__main__":
    main()
