// kadal_b200 — half-SM factorisation kernels (two CTAs per SM).
//
// The full-tile kernels of kb200_kernels.cu own an SM each (512 threads x 128 registers, 208 KiB
// of shared memory), so every phase of a CTA that is not tensor-pipe bound -- pipeline fill, the
// C tile fetch, the dependency chains of the triangular solve, the strictly sequential pivots of
// POTF2 (about 40 us per diagonal tile), the stores -- leaves the FP64 datapath of that SM idle
// (ncu r1i: 40-74 % DMMA pipe utilisation per launch).  Here every CTA is half an SM:
//
//   chol_panel_half_kernel  256 threads, 104 KiB: 64 rows x 128 columns of tile (i,j):
//                           X = (Psi_ij - L_i,0..j-1 L_j,0..j-1^T) L_jj^-T.  Rows of a triangular
//                           solve are independent, so the two halves of a tile never talk.
//                           (also prediction: 64 points x one sample tile)
//   chol_diag_half_kernel   256 threads, ~100 KiB: diagonal tile j: SYRK over the tile row
//                           (8 warps x 17 of the 136 lower 8x8 tiles: row blocks w and 15-w),
//                           fused forward substitution of [y F], blocked POTF2 on a PACKED lower
//                           triangle (66 KiB instead of 132), 8x8 inverse blocks, log-det.
//
// Two of them are resident per SM (of the same or of different launches / candidate groups), so the
// warp schedulers always have tensor-pipe work from one CTA while the other is in a latency-bound
// phase.  Same data layout, same arithmetic per element as the full-tile kernels (the GEMM core,
// the right-looking 8x8-inverse TRSM and the blocked POTF2 are the same algorithms).
// Reference call sites replaced: likelihood_func.py:168-213 (maincalc), prediction.py:245-251.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "kb200_kernels.h"

namespace kb200 {
namespace {

#ifndef KB200_POTF2_UNROLL    // warp POTF2: the rank-8 updates of the panels to the right unrolled over their blocks (A/B builds)
#define KB200_POTF2_UNROLL 1
#endif
#ifndef KB200_PANEL_PREFETCH  // panel CTAs prefetch the C half tile of the CTA this many blocks ahead into L2 (A/B builds).  Measured
#define KB200_PANEL_PREFETCH 0 // without gain at 296 / 592 (C2 7.80 / 7.81 vs 7.79 ms, C3 solve 7.41 vs 7.45 ms, C4 pred + s 8.62 vs 8.54): off
#endif
#ifndef KB200_RAGGED          // skip the identity padding of the ragged last tile (n % 128 != 0); 0: A/B builds
#define KB200_RAGGED 1
#endif
constexpr int HT = 256;                       // threads per CTA
constexpr int HWARPS = HT / 32;
constexpr int HROWS = 64;                     // rows of a half tile
constexpr int HALF_SLAB = HROWS * KS;         // 1024 doubles = 8 KiB
constexpr uint32_t HALF_BYTES = HALF_SLAB * 8;
constexpr int HNST = 4;                       // GEMM ring: stage = A half slab + B slab = 24 KiB
constexpr int HSTAGE = HALF_SLAB + SLAB_ELEMS;
constexpr int HLOOK = HNST - 2;
constexpr int DNST = 5;                       // SYRK ring of the diagonal kernel: one 16 KiB slab per stage
constexpr int DLOOK = DNST - 2;
constexpr int TLD = TB + 4;
constexpr unsigned FULLMASK = 0xffffffffu;

struct WarpPos {
  int wm, wn, g, t;
};

__device__ __forceinline__ int half_off(int r, int c) {   // element (r, c) of a 64 x 128 half tile in shared memory
  return (c >> 4) * HALF_SLAB + r * KS + ((((c >> 2) & 3) ^ (r & 3)) << 2) + (c & 3);
}
__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }   // packed lower triangle: S[tri(r) + c]

// C-fragment (c0,c1 = T[g][2t], T[g][2t+1]) of an 8x8 tile -> its two A-fragments (T[g][t], T[g][4+t])
__device__ __forceinline__ void cfrag_to_afrag(double c0, double c1, double& a0, double& a1, int lane) {
  const int src0 = (lane & ~3) + ((lane & 3) >> 1);
  const double v00 = __shfl_sync(FULLMASK, c0, src0), v01 = __shfl_sync(FULLMASK, c1, src0);
  const double v10 = __shfl_sync(FULLMASK, c0, src0 + 2), v11 = __shfl_sync(FULLMASK, c1, src0 + 2);
  a0 = (lane & 1) ? v01 : v00;
  a1 = (lane & 1) ? v11 : v10;
}

// sqrt(x), 1/(2 sqrt(x)) of a positive normal double (same sequence as the full-tile kernels)
__device__ __forceinline__ void sqrt_rsqrt(double x, double& s, double& h_out) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = x * y, h = 0.5 * y;
  double e = fma(-t, h, 0.5);
  t = fma(t, e, t);
  h = fma(h, e, h);
  e = fma(-t, h, 0.5);
  t = fma(t, e, t);
  h = fma(h, e, h);
  s = fma(fma(-t, t, x), h, t);
  h_out = h;
}

// ------------------------------------------------------------------ GEMM core (8 warps: 2 x 4, 4 x 4 blocks of 8 x 8 each)
// Block ownership is INTERLEAVED: warp (wm, wn) owns the row blocks 2 mt + wm and the column blocks 4 nt + wn
// (mt, nt = 0..3).  For a full tile this is the same work as contiguous 32 x 32 warp tiles (8 rows of a slab are 1 KiB
// apart whichever block they belong to: no change in bank conflicts).  For the RAGGED last tile of a matrix whose n is not
// a multiple of 128 the real row / column blocks 0 .. nrb-1 / 0 .. ncb-1 are spread evenly over the warps: the padded
// blocks (identity rows: exact zeros in every product) are skipped and the remaining work stays balanced -- n = 1000:
// the last half tile of every tile column has 40 real rows of 64 (5 blocks: 3 + 2 per warp instead of 4 + 4).
__device__ __forceinline__ int blk_row(const WarpPos& w, int mt) { return 8 * (2 * mt + w.wm); }
__device__ __forceinline__ int blk_col(const WarpPos& w, int nt) { return 8 * (4 * nt + w.wn); }

template <int NMT, int NNT>
__device__ __forceinline__ void mma_slab_h_t(double (&acc)[4][4][2], const double* As, const double* Bs, const WarpPos& w) {
#pragma unroll
  for (int kq = 0; kq < 4; ++kq) {
    const int sw = ((kq ^ (w.g & 3)) << 2) + w.t;
    double a[NMT], b[NNT];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) a[mt] = As[(blk_row(w, mt) + w.g) * KS + sw];
#pragma unroll
    for (int nt = 0; nt < NNT; ++nt) b[nt] = Bs[(blk_col(w, nt) + w.g) * KS + sw];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NNT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}
__device__ __forceinline__ void mma_slab_h(double (&acc)[4][4][2], const double* As, const double* Bs, const WarpPos& w) {
  mma_slab_h_t<4, 4>(acc, As, Bs, w);
}

// the same for a ragged tile: only the warp's first nmt row blocks and nnt column blocks are real (warp-uniform switch
// over compile-time shapes: predicating the 64 DMMAs of the full shape one by one measured slower than the padding it saves)
template <int NMT>
__device__ __forceinline__ void mma_slab_h_cols(double (&acc)[4][4][2], const double* As, const double* Bs, const WarpPos& w, int nnt) {
  switch (nnt) {
    case 4: mma_slab_h_t<NMT, 4>(acc, As, Bs, w); break;
    case 3: mma_slab_h_t<NMT, 3>(acc, As, Bs, w); break;
    case 2: mma_slab_h_t<NMT, 2>(acc, As, Bs, w); break;
    case 1: mma_slab_h_t<NMT, 1>(acc, As, Bs, w); break;
    default: break;
  }
}
__device__ __forceinline__ void mma_slab_h_part(double (&acc)[4][4][2], const double* As, const double* Bs, const WarpPos& w,
                                                int nmt, int nnt) {
  switch (nmt) {
    case 4: mma_slab_h_cols<4>(acc, As, Bs, w, nnt); break;
    case 3: mma_slab_h_cols<3>(acc, As, Bs, w, nnt); break;
    case 2: mma_slab_h_cols<2>(acc, As, Bs, w, nnt); break;
    case 1: mma_slab_h_cols<1>(acc, As, Bs, w, nnt); break;
    default: break;
  }
}

// ------------------------------------------------------------------ POTF2 on a packed lower triangle
// warp-level Cholesky of the 32x32 block at (o, o) (lane r owns row o + r); dinv[k] = 1 / L_kk.  Returns first bad pivot + 1 or 0.
//
// Blocked by 8 inside the warp: panel b (columns 8 b .. 8 b + 7 of the block) is factorised with the lane's 8 entries of it in
// registers -- 8 strictly sequential pivots, column k broadcast through lbuf (2 x 32 doubles) -- written back, and the panels to
// its right receive their rank-8 update from shared memory on DMMA (10 8x8 blocks per 32x32 block in all).  The loop over b is
// NOT unrolled: ~500 instructions that stay in the instruction cache.  (Round 2 first had all 32 pivots unrolled with the whole
// row in registers: ~3300 straight-line instructions run once per call by a single warp; ncu r2y attributed 38 % of that warp's
// stalls to instruction fetch (`no_inst`) and the call took ~10 us -- 64 % of the lifetime of a diagonal CTA at tile column 0
// with seven warps at the barrier.)
// The lane's own diagonal entry is tracked in a register of its own: its update by column k is -l * l with the lane's OWN l
// (the same fma, the same bits as a[lane] -= l * lb[lane]), so the chain from one pivot to the next is shuffle -> rsqrt +
// Newton -> l -> one fma, without the shared-memory round trip of the column broadcast.
__device__ __noinline__ int potf2_32_warp_t(double* S, int o, double* dinv, double* lbuf, int lane) {
  const int g = lane >> 2, tq = lane & 3;
  double* rowp = S + tri(o + lane) + o;       // element (o + lane, o)
  int bad = 0;
#pragma unroll 1
  for (int b = 0; b < 4; ++b) {
    const int c0 = 8 * b;
    const int rl = lane - c0;                 // row relative to the panel's diagonal block; < 0: row above the panel (idle)
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (rl >= c) ? rowp[c0 + c] : 0.0;
    double dg = (rl >= 0 && rl < 8) ? rowp[lane] : 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double piv = __shfl_sync(FULLMASK, dg, c0 + k);
      const bool ok = piv > 0.0;
      bad = (!ok && bad == 0) ? o + c0 + k + 1 : bad;
      piv = ok ? piv : 1.0;
      const double a2 = a[k] + a[k];
      double dk, h;
      sqrt_rsqrt(piv, dk, h);
      const double l = a2 * h;
      dg = fma(-l, l, dg);
      double* lb = lbuf + (k & 1) * 32;
      lb[lane] = l;
      __syncwarp();
#pragma unroll
      for (int c = k + 1; c < 8; ++c) a[c] = fma(-l, lb[c0 + c], a[c]);
      a[k] = (rl == k) ? dk : l;
      if (rl == k) dinv[o + c0 + k] = h + h;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (rl >= c) rowp[c0 + c] = a[c];
    __syncwarp();
    // rank-8 update of the panels to the right inside the block: C(q, p) -= L(q, b) L(p, b)^T, b < p <= q <= 3
#if KB200_POTF2_UNROLL
    // (unrolled over the 3 + 2 + 1 possible blocks, skipped by warp-uniform tests: the blocks of one column are independent and
    //  their loads / DMMAs overlap instead of running one block after the other)
#pragma unroll
    for (int p = 1; p < 4; ++p) {
      if (p <= b) continue;
      const double* br = S + tri(o + 8 * p + g) + o;
      const double b0 = br[c0 + tq], b1 = br[c0 + 4 + tq];
      const int cc = 8 * p + 2 * tq;
      double v0[3], v1[3], a0[3], a1[3];
#pragma unroll
      for (int q = p; q < 4; ++q) {
        const double* cr = S + tri(o + 8 * q + g) + o;
        v0[q - p] = cr[cc]; v1[q - p] = cr[cc + 1];
        a0[q - p] = -cr[c0 + tq]; a1[q - p] = -cr[c0 + 4 + tq];
      }
#pragma unroll
      for (int q = p; q < 4; ++q) dmma884(v0[q - p], v1[q - p], a0[q - p], b0);
#pragma unroll
      for (int q = p; q < 4; ++q) dmma884(v0[q - p], v1[q - p], a1[q - p], b1);
#pragma unroll
      for (int q = p; q < 4; ++q) {
        double* cr = S + tri(o + 8 * q + g) + o;
        const int r = 8 * q + g;
        if (cc <= r) cr[cc] = v0[q - p];
        if (cc + 1 <= r) cr[cc + 1] = v1[q - p];
      }
    }
#else
    for (int p = b + 1; p < 4; ++p) {
      const double* br = S + tri(o + 8 * p + g) + o;
      const double b0 = br[c0 + tq], b1 = br[c0 + 4 + tq];
      const int cc = 8 * p + 2 * tq;
      for (int q = p; q < 4; ++q) {
        double* cr = S + tri(o + 8 * q + g) + o;
        // (q == p: entries above the diagonal do not exist in the packed layout -- whatever is read there is not stored)
        double v0 = cr[cc], v1 = cr[cc + 1];
        dmma884(v0, v1, -cr[c0 + tq], b0);
        dmma884(v0, v1, -cr[c0 + 4 + tq], b1);
        const int r = 8 * q + g;
        if (cc <= r) cr[cc] = v0;
        if (cc + 1 <= r) cr[cc + 1] = v1;
      }
    }
#endif
    __syncwarp();
  }
  return bad;
}

// inverses of the four 8x8 diagonal sub-blocks of the 32x32 block at (o, o); out: 4 x 64 doubles
__device__ __forceinline__ void inv8_warp_t(const double* S, int o, const double* dinv, double* out, int lane) {
  const int blk = lane >> 3, c = lane & 7;
  const int r0 = o + 8 * blk;
  double w[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const double* Lr = S + tri(r0 + r) + r0;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < r; ++t) s = fma(Lr[t], w[t], s);
    w[r] = ((r == c) ? 1.0 : -s) * dinv[r0 + r];
    if (r < c) w[r] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) out[blk * 64 + r * 8 + c] = w[r];
}

// one 8-row strip against the 32x32 block at (o, o): rows <- rows * L_kk^-T (right-looking, DMMA).
// xrow: this lane's row (row g of the strip) at column o.
__device__ __forceinline__ void strip_trsm32_t(double* xrow, const double* S, int o, const double* W8, int lane) {
  const int g = lane >> 2, tq = lane & 3;
  double x[4][2];
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    x[ct][0] = xrow[8 * ct + 2 * tq];
    x[ct][1] = xrow[8 * ct + 2 * tq + 1];
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double a0, a1;
    cfrag_to_afrag(x[t][0], x[t][1], a0, a1, lane);
    double y0 = 0.0, y1 = 0.0;
    dmma884(y0, y1, a0, W8[t * 64 + g * 8 + tq]);
    dmma884(y0, y1, a1, W8[t * 64 + g * 8 + 4 + tq]);
    x[t][0] = y0;
    x[t][1] = y1;
    if (t < 3) {
      cfrag_to_afrag(y0, y1, a0, a1, lane);
      a0 = -a0;
      a1 = -a1;
#pragma unroll
      for (int ct = t + 1; ct < 4; ++ct) {
        const double* Lr = S + tri(o + 8 * ct + g) + o + 8 * t + tq;
        dmma884(x[ct][0], x[ct][1], a0, Lr[0]);
        dmma884(x[ct][0], x[ct][1], a1, Lr[4]);
      }
    }
  }
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    xrow[8 * ct + 2 * tq] = x[ct][0];
    xrow[8 * ct + 2 * tq + 1] = x[ct][1];
  }
}

// trailing update after the 32-wide panel at column o: items it0 <= it < it1 of "lower 16x16 blocks in
// row-major triangular order, then the RHS 8x16 blocks", warp `widx` of `nw`.
__device__ __forceinline__ void trailing_update32_t(double* S, double* trow, int o, int it0, int it1, int widx, int nw,
                                                    int lane, const int N = TB, const int tld = TLD) {
  const int g = lane >> 2, tq = lane & 3;
  const int base = o + 32;
  const int nb = (N - base) >> 4;
  const int ntri = nb * (nb + 1) / 2;
  const int nitems = ntri + nb;
  if (it1 > nitems) it1 = nitems;
  for (int it = it0 + widx; it < it1; it += nw) {
    if (it < ntri) {
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= it) ++bi;
      const int bj = it - bi * (bi + 1) / 2;
      double* ar0 = S + tri(base + 16 * bi + g);
      double* ar1 = S + tri(base + 16 * bi + g + 8);
      const double* br0 = S + tri(base + 16 * bj + g);
      const double* br1 = S + tri(base + 16 * bj + g + 8);
      const int cc = base + 16 * bj + 2 * tq;
      double c[2][2][2];
      // (bi == bj: the columns cc + 8.. of row block 0 lie above the diagonal -> not stored, not loaded)
      const bool up = bi > bj;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const bool ok0 = up || nt == 0;
        c[0][nt][0] = ok0 ? ar0[cc + 8 * nt] : 0.0;
        c[0][nt][1] = ok0 ? ar0[cc + 8 * nt + 1] : 0.0;
        c[1][nt][0] = ar1[cc + 8 * nt];
        c[1][nt][1] = ar1[cc + 8 * nt + 1];
      }
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) {
        const int kc = o + 4 * kq + tq;
        const double a0 = -ar0[kc], a1 = -ar1[kc];
        const double b0 = br0[kc], b1 = br1[kc];
        dmma884(c[0][0][0], c[0][0][1], a0, b0);
        dmma884(c[0][1][0], c[0][1][1], a0, b1);
        dmma884(c[1][0][0], c[1][0][1], a1, b0);
        dmma884(c[1][1][0], c[1][1][1], a1, b1);
      }
      // entries above the diagonal do not exist in the packed layout: store only c <= r
      const int r0 = base + 16 * bi + g, r1 = r0 + 8;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int c0 = cc + 8 * nt;
        if (c0 <= r0) ar0[c0] = c[0][nt][0];
        if (c0 + 1 <= r0) ar0[c0 + 1] = c[0][nt][1];
        if (c0 <= r1) ar1[c0] = c[1][nt][0];
        if (c0 + 1 <= r1) ar1[c0 + 1] = c[1][nt][1];
      }
    } else {
      const int bj = it - ntri;
      double* ar0 = trow + g * tld;
      const double* br0 = S + tri(base + 16 * bj + g);
      const double* br1 = S + tri(base + 16 * bj + g + 8);
      const int cc = base + 16 * bj + 2 * tq;
      double c[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { c[nt][0] = ar0[cc + 8 * nt]; c[nt][1] = ar0[cc + 8 * nt + 1]; }
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) {
        const int kc = o + 4 * kq + tq;
        const double a0 = -ar0[kc];
        dmma884(c[0][0], c[0][1], a0, br0[kc]);
        dmma884(c[1][0], c[1][1], a0, br1[kc]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { ar0[cc + 8 * nt] = c[nt][0]; ar0[cc + 8 * nt + 1] = c[nt][1]; }
    }
  }
}

// Blocked factorisation of the packed tile S with the RHS strip trow (8 x TLD).  All threads call.
__device__ void potf2_blocked_t(double* S, double* trow, double* dinv, double* W8s, double* lbuf, int* s_bad) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid == 0) *s_bad = 0;
  __syncthreads();
  for (int kb = 0; kb < 4; ++kb) {
    const int o = 32 * kb;
    if (warp == 0) {
      const int b = potf2_32_warp_t(S, o, dinv, lbuf, lane);
      __syncwarp();
      inv8_warp_t(S, o, dinv, W8s + kb * 256, lane);
      if (lane == 0 && b && *s_bad == 0) *s_bad = b;
    } else if (kb > 0) {
      trailing_update32_t(S, trow, o - 32, 3, 1 << 20, warp - 1, nwarp - 1, lane);
    }
    __syncthreads();
    const int nrow_strips = (TB - o - 32) >> 3;
    for (int st = warp; st <= nrow_strips; st += nwarp) {
      const int g = lane >> 2;
      if (st < nrow_strips) strip_trsm32_t(S + tri(o + 32 + 8 * st + g) + o, S, o, W8s + kb * 256, lane);
      else strip_trsm32_t(trow + g * TLD + o, S, o, W8s + kb * 256, lane);
    }
    __syncthreads();
    if (kb < 3) {
      trailing_update32_t(S, trow, o, 0, 3, warp, nwarp, lane);
      __syncthreads();
    }
  }
}

// accumulators of warp W of the diagonal kernel = minus its 17 tiles of Psi_jj (row blocks 15 - W and W)
template <int W>
__device__ __forceinline__ void syrk_acc_init(double (&acc)[17][2], const double* Cjj, int g, int t) {
  constexpr int NHI = 16 - W;
#pragma unroll
  for (int u = 0; u < 17; ++u) {
    const int R = u < NHI ? 15 - W : W, C = u < NHI ? u : u - NHI;
    const double2 p = *reinterpret_cast<const double2*>(Cjj + tile_off(8 * R + g, 8 * C + 2 * t));
    acc[u][0] = -p.x;
    acc[u][1] = -p.y;
  }
}

// one 16-wide slab of the SYRK for warp W: row blocks 15 - W (tiles C = 0 .. 15 - W) and W (C = 0 .. W).
// HI = false: the high row block is identity padding (ragged last tile of the matrix) and is skipped.
template <int W, bool HI = true>
__device__ __forceinline__ void syrk_slab(double (&acc)[17][2], const double* As, int g, int t) {
  constexpr int NHI = 16 - W;
#pragma unroll
  for (int kq = 0; kq < 4; ++kq) {
    const int sw = ((kq ^ (g & 3)) << 2) + t;
    const double a_hi = HI ? As[(8 * (15 - W) + g) * KS + sw] : 0.0, a_lo = As[(8 * W + g) * KS + sw];
#pragma unroll
    for (int u = HI ? 0 : NHI; u < 17; ++u) {
      const int C = u < NHI ? u : u - NHI;
      const double b = As[(8 * C + g) * KS + sw];
      dmma884(acc[u][0], acc[u][1], u < NHI ? a_hi : a_lo, b);
    }
  }
}

}  // namespace


// ============================================================================ small n: one CTA per matrix
// likelihood_func.py:93-102 + 168-213 for ONE candidate per CTA when the whole correlation matrix fits in shared
// memory (north star (2)): Psi + eps I is assembled straight into a PACKED lower
// triangle in shared memory (never touches HBM), factorised there by the same blocked POTF2 as a diagonal tile
// (32-wide panels: warp 0 factors the 32 x 32 block in registers with shuffles while warps 1..7 finish the previous
// trailing update; 8-row strips and the 16 x 16 trailing blocks on DMMA), with [y F] riding along as the right-hand-side
// strip (fused forward substitution); only Z = L^-1 [y F] (nq x n doubles), sum ln L_ii and info go back.
// N = n rounded up to 32 (identity padding, as the tile path pads to 128: n = 50 costs 64^3/3 flops instead of 128^3/3).
// Shared memory (doubles): S tri(N) | scratch max(8 (N + 4), d N + 32 lda + 2 d) (X staging during the assembly, then the
// RHS strip) | dinv N | W8 2 x 256 | lbuf 64 | red 16 | s_bad.
__host__ __device__ inline size_t lik_small_smem(int N, int d) {
  const size_t lda = (size_t)(d | 1);
  const size_t stage = (size_t)d * N + 32 * lda + 2 * (size_t)d + 8;
  const size_t strip = (size_t)8 * (N + 4);
  return ((size_t)N * (N + 1) / 2 + (stage > strip ? stage : strip) + N + 512 + 64 + 16 + 2) * 8;
}

// NTHR: 256 threads while two CTAs fit an SM (N <= 128), 512 beyond (one CTA per SM anyway: the strips and the trailing
// update of a panel then spread over 16 warps while warp 0 factors the next diagonal block)
template <bool FAST, int NTHR>
__global__ void __launch_bounds__(NTHR)
lik_small_kernel(CorrArgs C, CholArgs A, int N) {
  constexpr int NWARP = NTHR / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* S = reinterpret_cast<double*>(smem_raw);
  const int d = C.spec.d, lda = d | 1, n = C.n, tld = N + 4;
  double* scratch = S + (size_t)N * (N + 1) / 2;
  const size_t stage = (size_t)d * N + 32 * (size_t)lda + 2 * (size_t)d + 8, strip = (size_t)8 * tld;
  double* dinv = scratch + (stage > strip ? stage : strip);
  double* W8s = dinv + N;
  double* lbuf = W8s + 512;
  double* s_red = lbuf + 64;
  int* s_bad = reinterpret_cast<int*>(s_red + 16);
  // assembly staging inside `scratch`
  double* Bt = scratch;                         // [d][N] samples, transposed, zero padded
  double* As = Bt + (size_t)d * N;              // [32][lda] the rows of the current strip
  double* th2 = As + 32 * lda + ((32 * lda) & 1);   // [d] theta^2 | [d] RN(1/theta^2)   (16-byte aligned for double2 loads)
  double* trow = scratch;                       // afterwards: the RHS strip [8][tld]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long m = blockIdx.x;
  const double* cand = C.cand + m * C.cand_stride;
  const double w0 = cand[4 * C.spec.ntheta];
  const double eps = cand[4 * C.spec.ntheta + C.spec.nkernel];
  for (int idx = tid; idx < d * N; idx += NTHR) {
    const int k = idx / N, c = idx - k * N;
    Bt[idx] = c < n ? C.X[(size_t)c * d + k] : 0.0;
  }
  if (FAST && tid < 2 * d) th2[tid] = cand[2 * C.spec.ntheta + tid];
  // ---- assembly: 32-row strips, lane = row of the strip, warps walk the column quads 4 q <= row
  for (int r0 = 0; r0 < N; r0 += 32) {
    __syncthreads();                            // Bt / th2 ready; previous strip's As consumed
    for (int idx = tid; idx < 32 * d; idx += NTHR) {
      const int r = idx / d, k = idx - r * d;
      As[r * lda + k] = Bt[(size_t)k * N + r0 + r];
    }
    __syncthreads();
    const int row = r0 + lane;
    const int nq4 = (r0 + 32) >> 2;             // quads up to the end of the strip's diagonal block
    for (int q = warp; q < nq4; q += NWARP) {
      const int c0 = 4 * q;
      if (c0 > row) continue;                   // strictly above the diagonal: not stored
      Quad v;
      if (row >= n || c0 >= n) v = quad_set(0.0);
      else if (FAST) v = corr_quad_gauss<true, 0>(d, th2, th2 + d, w0, As + lane * lda, Bt + c0, N);
      else v = corr_quad<true>(C.spec, cand, As + lane * lda, Bt + c0, N);
      double* dst = S + tri(row) + c0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + e;
        if (c > row) break;
        double x = v.v[e];
        if (row >= n || c >= n) x = (row == c) ? 1.0 : 0.0;          // identity padding
        else if (row == c && C.add_eps) x = __dadd_rn(x, eps);       // Psi + eye * eps  (likelihood_func.py:171)
        dst[e] = x;
      }
    }
  }
  __syncthreads();                              // the staging area is dead: it becomes the RHS strip
  for (int idx = tid; idx < 8 * tld; idx += NTHR) {
    const int q = idx / tld, r = idx - q * tld;
    trow[idx] = (q < A.nq && r < n) ? A.R[(size_t)q * A.npad + r] : 0.0;
  }
  if (tid == 0) *s_bad = 0;
  __syncthreads();
  // ---- blocked factorisation with the RHS strip riding along (potf2_blocked_t for N / 32 panels)
  const int nblk = N >> 5;
  for (int kb = 0; kb < nblk; ++kb) {
    const int o = 32 * kb;
    double* W8k = W8s + (kb & 1) * 256;
    if (warp == 0) {
      const int b = potf2_32_warp_t(S, o, dinv, lbuf, lane);
      __syncwarp();
      inv8_warp_t(S, o, dinv, W8k, lane);
      if (lane == 0 && b && *s_bad == 0) *s_bad = b;
    } else if (kb > 0) {
      trailing_update32_t(S, trow, o - 32, 3, 1 << 20, warp - 1, NWARP - 1, lane, N, tld);
    }
    __syncthreads();
    const int nrow_strips = (N - o - 32) >> 3;
    for (int st = warp; st <= nrow_strips; st += NWARP) {
      const int g = lane >> 2;
      if (st < nrow_strips) strip_trsm32_t(S + tri(o + 32 + 8 * st + g) + o, S, o, W8k, lane);
      else strip_trsm32_t(trow + g * tld + o, S, o, W8k, lane);
    }
    __syncthreads();
    if (kb + 1 < nblk) {
      trailing_update32_t(S, trow, o, 0, 3, warp, NWARP, lane, N, tld);
      __syncthreads();
    }
  }
  // ---- outputs
  const int bad = *s_bad;
  if (tid == 0) A.info[m] = bad;
  double* z = A.Z + (size_t)m * A.nq * A.npad;
  for (int idx = tid; idx < A.nq * A.npad; idx += NTHR) {
    const int q = idx / A.npad, r = idx - q * A.npad;
    z[idx] = r < N ? trow[q * tld + r] : 0.0;
  }
  double v = 0.0;
  for (int i = tid; i < N; i += NTHR) v += log(fabs(S[tri(i) + i]));   // padded rows are exactly 1
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < NWARP; ++i) s += s_red[i];
    A.logdet[m] = s;
  }
}

// ============================================================================ diagonal tile
// smem (doubles): [0, DNST*2048) SYRK ring, aliased by the packed S (8256) | trow 8*TLD | dinv 128 |
//                 W8s 1024 | s_red 16 | lbuf 64 | barriers full[DNST], empty[DNST] | s_bad
constexpr int D_RING = DNST * SLAB_ELEMS;
constexpr int D_TROW = D_RING;
constexpr int D_DINV = D_TROW + 8 * TLD;
constexpr int D_W8 = D_DINV + TB;
constexpr int D_RED = D_W8 + 1024;
constexpr int D_LBUF = D_RED + 16;
constexpr int D_ZMAXQ = 4;                       // Z tiles are staged in shared memory for nq <= 4
constexpr int D_ZS = D_LBUF + 64;                // [2][D_ZMAXQ][128] rolling two-tile buffer
constexpr int D_BAR = D_ZS + 2 * D_ZMAXQ * TB;
constexpr size_t D_SMEM = (size_t)(D_BAR + 2 * DNST) * 8 + 16;
static_assert(TB * (TB + 1) / 2 <= D_RING, "packed S must fit in the ring");
static_assert(D_SMEM <= 113 * 1024, "two diagonal CTAs per SM");

__global__ void __launch_bounds__(HT, 2)
chol_diag_half_kernel(CholArgs A, int j) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* ring = sm;
  double* S = sm;
  double* trow = sm + D_TROW;
  double* dinv = sm + D_DINV;
  double* W8s = sm + D_W8;
  double* s_red = sm + D_RED;
  double* lbuf = sm + D_LBUF;
  double* zs = sm + D_ZS;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + D_BAR);
  uint64_t* empty = full + DNST;
  int* s_bad = reinterpret_cast<int*>(empty + DNST);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const long m = blockIdx.x;
  double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
  const double* rowj = tiles + (size_t)tri_index(j, 0) * TILE_ELEMS;
  double* Cjj = tiles + (size_t)tri_index(j, j) * TILE_ELEMS;
  const int nq = A.nq;
  const double* zb = A.Z + (size_t)m * A.nq * A.npad;
  if (tid == 0) {
    for (int i = 0; i < DNST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], HWARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  const int hiR = 15 - warp, loR = warp, nhi = hiR + 1;
  // ---- SYRK: warp w owns the 8-row blocks hiR = 15 - w (tiles C = 0..hiR) and loR = w (C = 0..w):
  //      17 of the 136 lower 8x8 tiles each.  acc[u]: u < nhi -> (hiR, u), else (loR, u - nhi).
  // The accumulators start at -Psi_jj (and -R_j), not at zero: the partial sums then walk from the entry down
  // to the (much smaller) Schur complement exactly as a right-looking update would, and every rounding is
  // relative to the current partial value instead of to the O(1) products that cancel at the end.  On
  // candidates with cond ~ 7e8 this takes the ln-likelihood error from 2e-7 to 2e-9 relative
  // (profiles/r1o_accumulate_from_c.md).
  double acc[17][2];
  switch (warp) {     // (a select between the two row blocks inside one unrolled loop crashes ptxas 12.9)
    case 0: syrk_acc_init<0>(acc, Cjj, g, t); break;
    case 1: syrk_acc_init<1>(acc, Cjj, g, t); break;
    case 2: syrk_acc_init<2>(acc, Cjj, g, t); break;
    case 3: syrk_acc_init<3>(acc, Cjj, g, t); break;
    case 4: syrk_acc_init<4>(acc, Cjj, g, t); break;
    case 5: syrk_acc_init<5>(acc, Cjj, g, t); break;
    case 6: syrk_acc_init<6>(acc, Cjj, g, t); break;
    default: syrk_acc_init<7>(acc, Cjj, g, t); break;
  }
  const int rr = tid >> 1, hh = tid & 1;     // RHS hook: row rr, columns 8 hh .. 8 hh + 7 of every slab
  double tacc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    tacc[q] = 0.0;
    if (q < nq && hh == 0)
      tacc[q] = -(A.rl_mode ? A.Rw[((size_t)m * A.nq + q) * A.npad + j * TB + rr] : A.R[(size_t)q * A.npad + j * TB + rr]);
  }

  const int nslab = A.rl_mode ? 0 : NSLAB * j;      // right-looking: tile (j, j) and the RHS are already updated
  // ragged last tile (n % 128 != 0): row blocks >= nrb are identity padding, their rows of L_j,0..j-1 are zero
  const int nrb = KB200_RAGGED ? (min(TB, A.n - j * TB) + 7) >> 3 : 16;
  const bool hi_on = hiR < nrb, lo_on = loR < nrb;
  auto issue = [&](int q) {
    const int st = q % DNST, use = q / DNST;
    if (use > 0) mbar_wait(&empty[st], (uint32_t)((use - 1) & 1));
    mbar_expect_tx(&full[st], SLAB_BYTES);
    bulk_g2s(ring + (size_t)st * SLAB_ELEMS, rowj + (size_t)q * SLAB_ELEMS, SLAB_BYTES, &full[st]);
  };
  if (tid == 0)
    for (int q = 0; q < DLOOK && q < nslab; ++q) issue(q);
  const bool zstaged = nq <= D_ZMAXQ;
  double zreg[2] = {0.0, 0.0};
  if (zstaged && nslab > 0) {
    for (int idx = tid; idx < nq * TB; idx += HT) zs[idx] = zb[(size_t)(idx >> 7) * A.npad + (idx & (TB - 1))];
    __syncthreads();
  }
#pragma unroll 1
  for (int q = 0; q < nslab; ++q) {
    const int st = q % DNST;
    if (tid == 0 && q + DLOOK < nslab) issue(q + DLOOK);
    mbar_wait(&full[st], (uint32_t)((q / DNST) & 1));
    const double* As = ring + (size_t)st * SLAB_ELEMS;
    if (hi_on) {
      switch (warp) {   // the warp's two row blocks are compile-time constants inside: no per-DMMA selects (ncu r2d: 128 FSEL per slab)
        case 0: syrk_slab<0>(acc, As, g, t); break;
        case 1: syrk_slab<1>(acc, As, g, t); break;
        case 2: syrk_slab<2>(acc, As, g, t); break;
        case 3: syrk_slab<3>(acc, As, g, t); break;
        case 4: syrk_slab<4>(acc, As, g, t); break;
        case 5: syrk_slab<5>(acc, As, g, t); break;
        case 6: syrk_slab<6>(acc, As, g, t); break;
        default: syrk_slab<7>(acc, As, g, t); break;
      }
    } else if (lo_on) {   // ragged last tile: the high row block of this warp is padding
      switch (warp) {
        case 0: syrk_slab<0, false>(acc, As, g, t); break;
        case 1: syrk_slab<1, false>(acc, As, g, t); break;
        case 2: syrk_slab<2, false>(acc, As, g, t); break;
        case 3: syrk_slab<3, false>(acc, As, g, t); break;
        case 4: syrk_slab<4, false>(acc, As, g, t); break;
        case 5: syrk_slab<5, false>(acc, As, g, t); break;
        case 6: syrk_slab<6, false>(acc, As, g, t); break;
        default: syrk_slab<7, false>(acc, As, g, t); break;
      }
    }
    // fused forward substitution: tacc[qq] += L_jk[rr][c] * Z_k[c]      (likelihood_func.py:183-193)
    // Z_k comes from a rolling two-tile shared-memory buffer: tile k + 1 is fetched into registers at the
    // first slab of tile k and stored at its last slab (the tile-boundary barrier publishes it), so the
    // L2 latency of Z never sits in front of a slab.  More than D_ZMAXQ right-hand sides: plain loads.
    {
      const int tile = q >> 3, sq = q & 7;
      const bool more = tile + 1 < j;
      if (zstaged && sq == 0 && more) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idx = tid + e * HT;
          if (idx < nq * TB) zreg[e] = zb[(size_t)(idx >> 7) * A.npad + (tile + 1) * TB + (idx & (TB - 1))];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int chunk = 2 * hh + cc;
        const double2* lp = reinterpret_cast<const double2*>(As + rr * KS + ((chunk ^ (rr & 3)) << 2));
        const double2 l01 = lp[0], l23 = lp[1];
#pragma unroll
        for (int qq = 0; qq < MAXQ; ++qq) {
          if (qq < nq) {
            double z0, z1, z2, z3;
            if (zstaged) {
              const double* z = zs + ((tile & 1) * nq + qq) * TB + sq * KS + 4 * chunk;
              const double2 za = *reinterpret_cast<const double2*>(z), zc = *reinterpret_cast<const double2*>(z + 2);
              z0 = za.x; z1 = za.y; z2 = zc.x; z3 = zc.y;
            } else {
              const double* z = zb + (size_t)qq * A.npad + q * KS + 4 * chunk;
              z0 = z[0]; z1 = z[1]; z2 = z[2]; z3 = z[3];
            }
            double sacc = tacc[qq];
            sacc = fma(l01.x, z0, sacc); sacc = fma(l01.y, z1, sacc);
            sacc = fma(l23.x, z2, sacc); sacc = fma(l23.y, z3, sacc);
            tacc[qq] = sacc;
          }
        }
      }
      if (zstaged && sq == 7 && more) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idx = tid + e * HT;
          if (idx < nq * TB) zs[(((tile + 1) & 1) * nq) * TB + idx] = zreg[e];
        }
      }
    }
    if (q > 0) {            // release one slab late (see chol_panel_half_kernel)
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[(q - 1) % DNST]);
    }
    if (zstaged && (q & 7) == 7) __syncthreads();   // tile boundary: publishes the staged Z tile
  }
  __syncthreads();        // every warp is done with the ring: S may overwrite it

  // ---- S = -acc = Psi_jj - sum (packed lower triangle), t = -tacc = R_j - sum
  for (int idx = tid; idx < 8 * TLD; idx += HT) trow[idx] = 0.0;
#pragma unroll
  for (int u = 0; u < 17; ++u) {
    const int R = u < nhi ? hiR : loR, C = u < nhi ? u : u - nhi;
    const int r = 8 * R + g, c = 8 * C + 2 * t;
    if (c <= r) S[tri(r) + c] = -acc[u][0];
    if (c + 1 <= r) S[tri(r) + c + 1] = -acc[u][1];
  }
  __syncthreads();
#pragma unroll
  for (int qq = 0; qq < MAXQ; ++qq) {
    if (qq < nq) {
      double s = tacc[qq];
      s += __shfl_xor_sync(FULLMASK, s, 1);
      if (hh == 0) trow[qq * TLD + rr] = -s;
    }
  }
  __syncthreads();
  potf2_blocked_t(S, trow, dinv, W8s, lbuf, s_bad);
  const int bad = *s_bad;
  if (tid == 0 && bad && A.info[m] == 0) A.info[m] = j * TB + bad;
  // ---- outputs: L_jj (tile layout, zeros above the diagonal), 8x8 inverse blocks, Z_j, log-det partial
  for (int idx = tid; idx < TILE_ELEMS / 4; idx += HT) {
    const int ch = idx & 3, row = (idx >> 2) & (TB - 1), s = idx >> 9;
    const int c0 = s * KS + 4 * ch;
    // quads strictly above the diagonal: zero since the assembly and never touched on the left-looking schedule (the
    // trailing updates of the right-looking one compute whole diagonal tiles: there the zeros are written back)
    if (c0 > row && !A.rl_mode) continue;
    double l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) l[e] = (c0 + e <= row) ? S[tri(row) + c0 + e] : 0.0;
    double2* dl = reinterpret_cast<double2*>(Cjj + s * SLAB_ELEMS + row * KS + ((ch ^ (row & 3)) << 2));
    dl[0] = make_double2(l[0], l[1]);
    dl[1] = make_double2(l[2], l[3]);
  }
  {
    double* w8 = A.W8 + ((size_t)m * A.nt + j) * 1024;
    for (int idx = tid; idx < 1024; idx += HT) w8[idx] = W8s[idx];
  }
  for (int idx = tid; idx < nq * TB; idx += HT) {
    const int q = idx >> 7, r = idx & (TB - 1);
    A.Z[((size_t)m * A.nq + q) * A.npad + j * TB + r] = trow[q * TLD + r];
  }
  {
    double v = (tid < TB) ? log(fabs(S[tri(tid) + tid])) : 0.0;   // padded rows are exactly 1
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < 4; ++i) s += s_red[i];
      A.logdet[m] = (j == 0 ? 0.0 : A.logdet[m]) + s;
    }
  }
}

// ============================================================================ panel (half tile)
// smem (doubles): [0, HNST*HSTAGE) GEMM ring (4 x 24 KiB); afterwards Cs (64 x 128) in [0, 8192) and the
//   slabs 0, 1 of L_jj in [8192, 12288); once the rows are in registers Cs is dead and the slabs 2..7 of
//   L_jj -- only their rows >= 16 s, which is all the solve reads -- are packed into [0, 6912).  The
//   whole factor tile is then resident: one single-use mbarrier per slab, no slot is ever re-filled
//   while the solve runs and the eight warps drift freely.  W8s at 12288; barriers behind it.
constexpr int P_CS = HROWS * TB;                       // 8192
constexpr int P_WS01 = P_CS;                           // slabs 0 and 1, complete
constexpr int P_W8 = HNST * HSTAGE;                    // 12288
constexpr int P_BAR = P_W8 + 1024;
constexpr int P_NBAR = 2 * HNST + NSLAB + 1;           // full, empty, one per L_jj slab, W8
constexpr size_t P_SMEM = (size_t)(P_BAR + P_NBAR) * 8 + 16;
static_assert(P_WS01 + 2 * SLAB_ELEMS <= P_W8, "slabs 0, 1 must fit behind Cs");
static_assert(P_SMEM <= 113 * 1024, "two panel CTAs per SM");
__host__ __device__ constexpr int wslab_rows(int s) { return TB - KS * s; }           // rows 16 s .. 127
__host__ __device__ constexpr int wslab_base(int s) {                                  // packed offset, s >= 2
  int o = 0;
  for (int q = 2; q < s; ++q) o += wslab_rows(q) * KS;
  return o;
}
static_assert(wslab_base(NSLAB) <= P_CS, "trimmed slabs 2..7 must fit in the dead Cs area");

// the right-looking solve of one warp's 8 rows (C-fragment layout, 16 column tiles) against the resident L_jj:
// finished tile x 8x8 inverse = 2 DMMA, then DMMA updates to its right.  RAGGED: only the first ncb column blocks are real.
template <bool RAGGED>
__device__ __forceinline__ void panel_solve_rows(double (&x)[16][2], const double* stages, const double* W8s, uint64_t* wbar,
                                                 int lane, int ncb) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int s = 0; s < NSLAB; ++s) {
    if (RAGGED && 2 * s >= ncb) break;
    mbar_wait(&wbar[s], 0);
    // virtual slab base: row r of slab s at Ls[r * KS + ...] (rows < 16 s of the trimmed slabs are not there)
    const double* Ls = s < 2 ? stages + P_WS01 + s * SLAB_ELEMS : stages + wslab_base(s) - KS * s * KS;
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
      const int t = 2 * s + tt;
      if (RAGGED && t >= ncb) break;
      double a0, a1;
      cfrag_to_afrag(x[t][0], x[t][1], a0, a1, lane);
      double y0 = 0.0, y1 = 0.0;
      dmma884(y0, y1, a0, W8s[t * 64 + g * 8 + tq]);
      dmma884(y0, y1, a1, W8s[t * 64 + g * 8 + 4 + tq]);
      x[t][0] = y0;
      x[t][1] = y1;
      if (t < 15) {
        cfrag_to_afrag(y0, y1, a0, a1, lane);
        a0 = -a0;
        a1 = -a1;
#pragma unroll
        for (int ct = t + 1; ct < 16; ++ct) {
          if (RAGGED && ct >= ncb) break;
          const double* lrow = Ls + (8 * ct + g) * KS + tq;
          const double b0 = lrow[((2 * tt) ^ (g & 3)) << 2];
          const double b1 = lrow[((2 * tt + 1) ^ (g & 3)) << 2];
          dmma884(x[ct][0], x[ct][1], a0, b0);
          dmma884(x[ct][0], x[ct][1], a1, b1);
        }
      }
    }
  }
}

struct PanelPtrs {
  const double *rowA, *rowB, *Ljj, *W8j;
  double* Cij;
  long ssq_base;
  int hr, rows_real, cols_real;
  int solve_after;       // prediction, right-looking, fused sweep (rl_mode 4): this CTA's tile is final after its update: solve it
};

template <bool RAGGED>
__device__ __forceinline__ void panel_half_body(const CholArgs& A, int j, const PanelPtrs& P, unsigned char* smem_raw) {
  double* stages = reinterpret_cast<double*>(smem_raw);
  double* Cs = stages;
  double* W8s = stages + P_W8;
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + P_BAR);
  uint64_t* empty = full + HNST;
  uint64_t* wbar = empty + HNST;        // [NSLAB]
  uint64_t* w8bar = wbar + NSLAB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const WarpPos w{warp & 1, warp >> 1, lane >> 2, lane & 3};
  const double *rowA = P.rowA, *rowB = P.rowB, *Ljj = P.Ljj, *W8j = P.W8j;
  double* Cij = P.Cij;
  const long ssq_base = P.ssq_base;
  const int hr = P.hr, rows_real = P.rows_real, cols_real = P.cols_real;
  const int nrb = (rows_real + 7) >> 3, ncb = (cols_real + 7) >> 3;      // real 8-row / 8-column blocks: 1..8, 1..16
  const int nmt = (nrb - w.wm + 1) >> 1, nnt = (ncb - w.wn + 3) >> 2;    // ... of this warp (blocks 2 mt + wm, 4 nt + wn)
  constexpr bool ragged = RAGGED;                    // CTA-uniform: nrb < 8 || ncb < 16
  const double* gA = rowA + hr * HALF_SLAB;          // rows 64 hr .. of every slab are 8 KiB contiguous
  if (tid == 0) {
    for (int i = 0; i < HNST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], HWARPS); }
    for (int i = 0; i < NSLAB; ++i) mbar_init(&wbar[i], 1);
    mbar_init(w8bar, 1);
    fence_mbar_init();
    mbar_expect_tx(w8bar, 8192);
    bulk_g2s(W8s, W8j, 8192, w8bar);
  }
  __syncthreads();

  // ---- GEMM: acc = -C + A(64 x 128 j) B(128 x 128 j)^T, operands streamed slab by slab
  // (right-looking variant: no accumulation before the solve (T), one tile of depth for the update (U))
  const int nslab = A.rl_mode == 1 ? 0 : (A.rl_mode >= 2 ? NSLAB : NSLAB * j);
  auto issue = [&](int q) {
    const int st = q % HNST, use = q / HNST;
    if (use > 0) mbar_wait(&empty[st], (uint32_t)((use - 1) & 1));
    double* dst = stages + (size_t)st * HSTAGE;
    mbar_expect_tx(&full[st], HALF_BYTES + SLAB_BYTES);
    bulk_g2s(dst, gA + (size_t)q * SLAB_ELEMS, HALF_BYTES, &full[st]);
    bulk_g2s(dst + HALF_SLAB, rowB + (size_t)q * SLAB_ELEMS, SLAB_BYTES, &full[st]);
  };
  // (the first slabs are in flight while the C tile is fetched)
  if (tid == 0)
    for (int q = 0; q < HLOOK && q < nslab; ++q) issue(q);

  // ---- every warp takes 8 complete rows (C-fragment layout, 16 column tiles) and solves them
  const int g = w.g, tq = w.t;
  const int xr = 8 * warp + g;
  double x[16][2];
  if (nslab == 0) {
    // pure solve (tile column 0; every solve of the right-looking schedule): nothing to accumulate, so no detour through
    // the accumulators and shared memory -- the rows come straight from the tile, the whole of L_jj is requested at once
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const uint32_t bytes = s < 2 ? SLAB_BYTES : (uint32_t)wslab_rows(s) * KS * 8;
        mbar_expect_tx(&wbar[s], bytes);
        if (s < 2) bulk_g2s(stages + P_WS01 + s * SLAB_ELEMS, Ljj + (size_t)s * SLAB_ELEMS, bytes, &wbar[s]);
        else bulk_g2s(stages + wslab_base(s), Ljj + (size_t)s * SLAB_ELEMS + (size_t)KS * s * KS, bytes, &wbar[s]);
      }
    }
#pragma unroll
    for (int ct = 0; ct < 16; ++ct) {
      const double2 p = *reinterpret_cast<const double2*>(Cij + tile_off(HROWS * hr + xr, 8 * ct + 2 * tq));
      x[ct][0] = p.x;
      x[ct][1] = p.y;
    }
  } else {
    // acc starts at -C_ij (see chol_diag_half_kernel: rounding relative to the shrinking partial value)
    double acc[4][4][2];
  #pragma unroll
    for (int mt = 0; mt < 4; ++mt)
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = blk_row(w, mt) + w.g, c = blk_col(w, nt) + 2 * w.t;
        double2 p = make_double2(0.0, 0.0);
        if (!ragged || (mt < nmt && nt < nnt)) p = *reinterpret_cast<const double2*>(Cij + tile_off(HROWS * hr + r, c));
        acc[mt][nt][0] = -p.x;
        acc[mt][nt][1] = -p.y;
      }

  #pragma unroll 1
    for (int q = 0; q < nslab; ++q) {
      const int st = q % HNST;
      if (tid == 0 && q + HLOOK < nslab) issue(q + HLOOK);
      mbar_wait(&full[st], (uint32_t)((q / HNST) & 1));
      const double* As = stages + (size_t)st * HSTAGE;
      if (!ragged) mma_slab_h(acc, As, As + HALF_SLAB, w);
      else mma_slab_h_part(acc, As, As + HALF_SLAB, w, nmt, nnt);
      // Consumer release, one slab late: the stage of slab q - 1 is handed back here.  The bulk copy that
      // re-fills a stage runs in the async proxy and is not ordered behind generic-proxy loads that are
      // merely *issued* (an arrive right behind the last LDS of slab q let the copy land under pending loads
      // with two CTAs per SM, profiles/r1k_half_race.md).  Every DMMA of slab q - 1 precedes this point in
      // the instruction stream, and a DMMA cannot issue before the loads of its operands have returned, so
      // nothing of slab q - 1 is still being read.  The ring has a slab of slack for it (refill = slab q - 2).
      if (q > 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[(q - 1) % HNST]);
      }
    }
    if (A.rl_mode >= 2 && !P.solve_after) {   // trailing update: C -= acc in place, nothing to solve
  #pragma unroll
      for (int mt = 0; mt < 4; ++mt)
  #pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r = blk_row(w, mt) + w.g, c = blk_col(w, nt) + 2 * w.t;
          if (!ragged || (mt < nmt && nt < nnt))
            *reinterpret_cast<double2*>(Cij + tile_off(HROWS * hr + r, c)) = make_double2(-acc[mt][nt][0], -acc[mt][nt][1]);
        }
      return;
    }
    fence_proxy_async();
    __syncthreads();       // ring drained: Cs and the first two L_jj slabs may overwrite it
    if (tid == 0) {
  #pragma unroll
      for (int s = 0; s < 2; ++s) {
        mbar_expect_tx(&wbar[s], SLAB_BYTES);
        bulk_g2s(stages + P_WS01 + s * SLAB_ELEMS, Ljj + (size_t)s * SLAB_ELEMS, SLAB_BYTES, &wbar[s]);
      }
    }
    // ---- C = -acc = Psi_ij - sum -> Cs
  #pragma unroll
    for (int mt = 0; mt < 4; ++mt)
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = blk_row(w, mt) + w.g, c = blk_col(w, nt) + 2 * w.t;
        *reinterpret_cast<double2*>(Cs + half_off(r, c)) = make_double2(-acc[mt][nt][0], -acc[mt][nt][1]);
      }
    __syncthreads();
  #pragma unroll
    for (int ct = 0; ct < 16; ++ct) {
      const double2 p = *reinterpret_cast<const double2*>(Cs + half_off(xr, 8 * ct + 2 * tq));
      x[ct][0] = p.x;
      x[ct][1] = p.y;
    }
    // Cs is dead: the rest of L_jj (rows >= 16 s of slab s) goes where it was
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
  #pragma unroll
      for (int s = 2; s < NSLAB; ++s) {
        const uint32_t bytes = (uint32_t)wslab_rows(s) * KS * 8;
        mbar_expect_tx(&wbar[s], bytes);
        bulk_g2s(stages + wslab_base(s), Ljj + (size_t)s * SLAB_ELEMS + (size_t)KS * s * KS, bytes, &wbar[s]);
      }
    }
  }
  mbar_wait(w8bar, 0);
  if (!ragged) {
    panel_solve_rows<false>(x, stages, W8s, wbar, lane, 16);
  } else {
    if (8 * warp < rows_real) {
      if (ncb == 16) panel_solve_rows<false>(x, stages, W8s, wbar, lane, 16);     // ragged rows only: the plain sweep
      else panel_solve_rows<true>(x, stages, W8s, wbar, lane, ncb);
    }
    // (a ragged solve does not wait for the slabs it does not read: every copy has to land before the CTA may exit)
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) mbar_wait(&wbar[s], 0);
  }
  // ---- store X and, for prediction, accumulate ||x_row||^2
  if (ragged && 8 * warp >= rows_real) return;     // padded rows: the zeros in place stay (likelihood; points are never padded)
  double rs = 0.0;
#pragma unroll
  for (int ct = 0; ct < 16; ++ct) {
    if (ragged && ct >= ncb) break;                // padded columns: zeros in place
    *reinterpret_cast<double2*>(Cij + tile_off(HROWS * hr + xr, 8 * ct + 2 * tq)) = make_double2(x[ct][0], x[ct][1]);
    rs = fma(x[ct][0], x[ct][0], rs);
    rs = fma(x[ct][1], x[ct][1], rs);
  }
  if (A.pred_mode) {
    rs += __shfl_xor_sync(FULLMASK, rs, 1);
    rs += __shfl_xor_sync(FULLMASK, rs, 2);
    if (tq == 0) {
      double* dst = A.ssq + ssq_base + xr;
      *dst = (j == 0 && !P.solve_after ? 0.0 : *dst) + rs;
    }
  }
}

__global__ void __launch_bounds__(HT, 2)
chol_panel_half_kernel(CholArgs A, int j) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int hr = blockIdx.x & 1;                     // which 64 rows of the tile
  const double *rowA, *rowB, *Ljj, *W8j;
  double* Cij;
  long ssq_base = 0;
  // real rows of this half tile / real columns of the output tile (ragged last tile of a matrix with n % 128 != 0: the
  // rest is identity padding -- zero rows of L, zero columns of psi -- and is skipped, see mma_slab_h)
  int rows_real = HROWS, cols_real = TB, solve_after = 0;
  if (A.pred_mode && A.rl_mode >= 2) {
    // prediction, right-looking (few point tiles, many sample tiles): after V_j = (psi_j - ...) L_jj^-T every later tile of
    // the point-tile row is updated at once, psi_i -= V_j L_ij^T for i > j: (nt - 1 - j) x more CTAs than the left-looking
    // launch of column j has.  blockIdx.x >> 1 = pt (nt - 1 - j) + a, i = j + 1 + a.
    const int M = A.nt - 1 - j;
    const long t = blockIdx.x >> 1, pt = t / M;
    const int i = j + 1 + (int)(t - pt * M);
    double* base = A.chunk + (size_t)pt * A.nt * TILE_ELEMS;
    rowA = base + (size_t)j * TILE_ELEMS;                            // V_j, 8 slabs
    rowB = A.L + (size_t)tri_index(i, j) * TILE_ELEMS;               // L_ij, 8 slabs
    Cij = base + (size_t)i * TILE_ELEMS;
    Ljj = A.L;                                                       // unused
    W8j = A.W8 + (size_t)j * 1024;                                   // fetched, unused
    cols_real = min(TB, A.n - i * TB);
    if (A.rl_mode == 4 && i == j + 1) {
      // fused sweep: sample tile j + 1 is final once column j has updated it -- this CTA goes on to solve it (V_{j+1}) instead of
      // leaving that to a launch of its own: one launch per tile column instead of two in a chain of nt columns
      solve_after = 1;
      Ljj = A.L + (size_t)tri_index(i, i) * TILE_ELEMS;
      W8j = A.W8 + (size_t)i * 1024;
      ssq_base = pt * TB + hr * HROWS;
    }
  } else if (A.pred_mode) {
    const long pt = blockIdx.x >> 1;
    double* base = A.chunk + (size_t)pt * A.nt * TILE_ELEMS;
    rowA = base;
    Cij = base + (size_t)j * TILE_ELEMS;
    rowB = A.L + (size_t)tri_index(j, 0) * TILE_ELEMS;
    Ljj = A.L + (size_t)tri_index(j, j) * TILE_ELEMS;
    W8j = A.W8 + (size_t)j * 1024;
    ssq_base = pt * TB + hr * HROWS;
    cols_real = min(TB, A.n - j * TB);
  } else if (A.rl_mode >= 2) {
    // right-looking trailing update after column k = j: C_ib -= L_ik L_bk^T for k < b <= i.
    // mode 2: all pairs from row/column k + 1 + i0 on: blockIdx.x >> 1 = a (a + 1) / 2 + c enumerates
    //         (i, b) = (k + 1 + i0 + a, k + 1 + i0 + c), c <= a;  mode 3: the strip b = k + 1 only (look-ahead)
    const int t = blockIdx.x >> 1;
    int a, c;
    if (A.rl_mode == 3) {
      a = t;
      c = 0;
    } else {
      a = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
      while ((a + 1) * (a + 2) / 2 <= t) ++a;
      while (a * (a + 1) / 2 > t) --a;
      c = t - a * (a + 1) / 2;
    }
    const int i = j + 1 + A.i0 + a, b = j + 1 + A.i0 + c;
    const long m = blockIdx.y;
    double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
    rowA = tiles + (size_t)tri_index(i, j) * TILE_ELEMS;       // the 8 slabs of tile (i, k)
    rowB = tiles + (size_t)tri_index(b, j) * TILE_ELEMS;       // the 8 slabs of tile (b, k)
    Cij = tiles + (size_t)tri_index(i, b) * TILE_ELEMS;
    Ljj = tiles;                                               // unused
    W8j = A.W8 + ((size_t)m * A.nt + j) * 1024;                // fetched, unused
    rows_real = min(HROWS, A.n - i * TB - hr * HROWS);
    cols_real = min(TB, A.n - b * TB);
  } else {
    const int i = j + 1 + A.i0 + (blockIdx.x >> 1);
    const long m = blockIdx.y;
    double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
    rowA = tiles + (size_t)tri_index(i, 0) * TILE_ELEMS;
    rowB = tiles + (size_t)tri_index(j, 0) * TILE_ELEMS;
    Cij = tiles + (size_t)tri_index(i, j) * TILE_ELEMS;
    Ljj = tiles + (size_t)tri_index(j, j) * TILE_ELEMS;
    W8j = A.W8 + ((size_t)m * A.nt + j) * 1024;
    rows_real = min(HROWS, A.n - i * TB - hr * HROWS);
  }
#if !KB200_RAGGED
  rows_real = HROWS;
  cols_real = TB;
#endif
#if KB200_PANEL_PREFETCH
  // Left-looking launches: pull the C half tile of the CTA that will take this slot about one wave later (CTAs start in
  // block order, 2 x 148 are resident) into L2.  ncu r2zz: a prediction CTA of sample tile 1 spent 20 % of its life waiting
  // for its 64 x 128 rows of psi to come from HBM before the first DMMA could issue.
  if (A.rl_mode == 0) {
    const long lin = (long)blockIdx.y * gridDim.x + blockIdx.x + KB200_PANEL_PREFETCH;
    if (lin < (long)gridDim.x * gridDim.y) {
      const int bx = (int)(lin % gridDim.x);
      const long by = lin / gridDim.x;
      const double* Cn = A.pred_mode
                             ? A.chunk + ((size_t)(bx >> 1) * A.nt + j) * TILE_ELEMS
                             : A.L + ((size_t)by * A.tiles_per_mat + tri_index(j + 1 + A.i0 + (bx >> 1), j)) * TILE_ELEMS;
      const int nsl = (cols_real + KS - 1) / KS;       // slabs with real columns (the same tile column: the same count)
      for (int line = threadIdx.x; line < nsl * HROWS; line += HT) {   // one 128-byte line = one row of one slab
        const double* q = Cn + (size_t)(line >> 6) * SLAB_ELEMS + (size_t)(HROWS * (bx & 1) + (line & 63)) * KS;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
      }
    }
  }
#endif
  if (rows_real <= 0) return;                        // a half tile of padding only: zeros stay zeros (whole CTA, before any barrier)
  const PanelPtrs P{rowA, rowB, Ljj, W8j, Cij, ssq_base, hr, rows_real, cols_real, solve_after};
  // CTA-uniform dispatch: the full-tile body is compiled without any trace of the ragged one
  if (rows_real < HROWS || cols_real < TB) panel_half_body<true>(A, j, P, smem_raw);
  else panel_half_body<false>(A, j, P, smem_raw);
}

// ============================================================================ pure solves
// X = C L_jj^-T with nothing to accumulate first: tile column 0 of the left-looking schedule, every solve of the
// right-looking one, sample tile 0 (or every sample tile, right-looking) of a prediction.  chol_panel_half_kernel spends more
// than half of such a CTA's life outside the tensor pipe (fetching the 64 x 128 rows, fetching 94 KiB of L_jj and W8 for
// 1 Mflop of work, storing; ncu r2y: 56 % DMMA pipe).  Here L_jj and its 8x8 inverses are fetched ONCE per CTA and stay
// resident while the CTA walks a list of half tiles that share them (all point tiles of a prediction; the tiles below the
// diagonal tile of one candidate); there is no CTA-wide barrier after that, every warp streams 8-row strips
// (load -> 272 DMMA -> store) at its own pace, so the sixteen warps of an SM drift out of phase and the pipe always has a
// few strips in their DMMA-rich early steps.  Same arithmetic per row as the panel kernel (panel_solve_rows): bit-identical.
// grid: (CTAs per matrix, matrices); half tile h of the matrix: tile j + 1 + i0 + (h >> 1) (likelihood) / point tile h >> 1.
// NTHR: 256 (two CTAs per SM, default) or one wide CTA per SM ($KB200_PURE_SOLVE_THREADS=640: 20 strips in flight per copy of
// L_jj at 96 registers; measured SLOWER -- C3 variance solve 7.78 vs 7.44 ms per 4e6 points, 768 threads 7.83: more warps do
// not help, the strips are bound by the pipe and by their 16 KB of HBM traffic each, not by latency).
template <int NTHR>
__global__ void __launch_bounds__(NTHR, NTHR <= 256 ? 2 : 1)
solve_rows_half_kernel(CholArgs A, int j, int nhalf) {
  constexpr int NW = NTHR / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);
  double* W8s = stages + P_W8;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(stages + P_BAR);   // [NSLAB]
  uint64_t* w8bar = wbar + NSLAB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
  const long m = blockIdx.y;
  double* tiles = A.L + (A.pred_mode ? 0 : (size_t)m * A.tiles_per_mat * TILE_ELEMS);
  const double* Ljj = tiles + (size_t)tri_index(j, j) * TILE_ELEMS;
  const double* W8j = A.W8 + ((A.pred_mode ? 0 : (size_t)m * A.nt) + j) * 1024;
  if (tid == 0) {
    for (int i = 0; i < NSLAB; ++i) mbar_init(&wbar[i], 1);
    mbar_init(w8bar, 1);
    fence_mbar_init();
    mbar_expect_tx(w8bar, 8192);
    bulk_g2s(W8s, W8j, 8192, w8bar);
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) {
      const uint32_t bytes = s < 2 ? SLAB_BYTES : (uint32_t)wslab_rows(s) * KS * 8;
      mbar_expect_tx(&wbar[s], bytes);
      if (s < 2) bulk_g2s(stages + P_WS01 + s * SLAB_ELEMS, Ljj + (size_t)s * SLAB_ELEMS, bytes, &wbar[s]);
      else bulk_g2s(stages + wslab_base(s), Ljj + (size_t)s * SLAB_ELEMS + (size_t)KS * s * KS, bytes, &wbar[s]);
    }
  }
  __syncthreads();
  const int cols_real = A.pred_mode && KB200_RAGGED ? min(TB, A.n - j * TB) : TB;
  const int ncb = (cols_real + 7) >> 3;
  // The CTA's half tiles are blockIdx.x, blockIdx.x + gridDim.x, ...; its strips (8 rows each, 8 per half tile) are numbered
  // through them and dealt to the warps round robin: strip k = rows 8 (k & 7) .. of half tile blockIdx.x + (k >> 3) gridDim.x.
  const int nh = blockIdx.x < nhalf ? (nhalf - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int nstrip = 8 * nh;
  // first element (row g of the strip, columns 2 tq ..) of strip k in its tile, or null: padded rows; rb = row of the half tile
  auto strip = [&](int k, int& t, int& rb) -> double* {
    const int h = blockIdx.x + (k >> 3) * gridDim.x, hr = h & 1;
    t = h >> 1;
    rb = HROWS * hr + 8 * (k & 7);
    if (A.pred_mode) return A.chunk + ((size_t)t * A.nt + j) * TILE_ELEMS + tile_off(rb + g, 2 * tq);
    const int i = j + 1 + A.i0 + t;
    if (KB200_RAGGED && rb >= A.n - i * TB) return nullptr;
    return tiles + (size_t)tri_index(i, j) * TILE_ELEMS + tile_off(rb + g, 2 * tq);
  };
  auto next_strip = [&](int& k, int& t, int& rb) -> double* {  // advances k to the warp's next real strip
    for (k += NW; k < nstrip; k += NW) {
      double* p = strip(k, t, rb);
      if (p) return p;
    }
    return nullptr;
  };
  auto load = [&](double (&x)[16][2], const double* p0, int rb) {
    const double* base = p0 - tile_off(rb + g, 2 * tq);
#pragma unroll
    for (int ct = 0; ct < 16; ++ct) {
      const double2 p = *reinterpret_cast<const double2*>(base + tile_off(rb + g, 8 * ct + 2 * tq));
      x[ct][0] = p.x;
      x[ct][1] = p.y;
    }
  };
  auto finish = [&](double (&x)[16][2], double* p0, int rb, int t) {
    if (ncb == 16) panel_solve_rows<false>(x, stages, W8s, wbar, lane, 16);
    else panel_solve_rows<true>(x, stages, W8s, wbar, lane, ncb);
    double* base = p0 - tile_off(rb + g, 2 * tq);
    double rs = 0.0;
#pragma unroll
    for (int ct = 0; ct < 16; ++ct) {
      if (ct >= ncb) break;                          // padded columns: zeros in place
      *reinterpret_cast<double2*>(base + tile_off(rb + g, 8 * ct + 2 * tq)) = make_double2(x[ct][0], x[ct][1]);
      rs = fma(x[ct][0], x[ct][0], rs);
      rs = fma(x[ct][1], x[ct][1], rs);
    }
    if (A.pred_mode) {
      rs += __shfl_xor_sync(FULLMASK, rs, 1);
      rs += __shfl_xor_sync(FULLMASK, rs, 2);
      if (tq == 0) {
        double* dst = A.ssq + (long)t * TB + rb + g;
        *dst = (j == 0 ? 0.0 : *dst) + rs;
      }
    }
  };
  // (a second register set for the next strip's rows spills at the register cap: the next strip is pulled into L2 instead --
  //  its rows come from HBM, written by the psi kernel / the assembly long before)
  auto prefetch = [&](const double* p0, int rb) {
    const char* tile0 = reinterpret_cast<const char*>(p0 - tile_off(rb + g, 2 * tq));
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int line = lane + 32 * e;                // 64 lines of 128 B: slab line >> 3, row rb + (line & 7)
      const char* q = tile0 + (size_t)(line >> 3) * SLAB_BYTES + (size_t)(rb + (line & 7)) * (KS * 8);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
  };
  double x[16][2];
  int k = warp - NW, t = 0, rb = 0;
  double* p = next_strip(k, t, rb);
  bool first = true;
  while (p) {
    const int tc = t, rc = rb;
    load(x, p, rc);
    double* pn = next_strip(k, t, rb);
    if (pn) prefetch(pn, rb);
    if (first) {
      mbar_wait(w8bar, 0);
      first = false;
    }
    finish(x, p, rc, tc);
    p = pn;
  }
  // every copy has to land before the CTA may exit (a warp without a real strip, a ragged solve that stops early)
  mbar_wait(w8bar, 0);
#pragma unroll
  for (int s = 0; s < NSLAB; ++s) mbar_wait(&wbar[s], 0);
}

// right-looking RHS update after column k: rw_i -= L_ik z_k for the tiles i > k (nq right-hand sides);
// one CTA per (i, candidate), thread = row; column k == -1 initialises rw = R
__global__ void __launch_bounds__(TB)
rl_rhs_update_kernel(CholArgs A, int k) {
  const long m = blockIdx.y;
  double* rw = A.Rw + (size_t)m * A.nq * A.npad;
  if (k < 0) {
    for (int idx = blockIdx.x * TB + threadIdx.x; idx < A.nq * A.npad; idx += gridDim.x * TB) rw[idx] = A.R[idx];
    return;
  }
  __shared__ double z[MAXQ][TB];
  const int i = k + 1 + A.i0 + blockIdx.x, r = threadIdx.x;
  const double* zk = A.Z + (size_t)m * A.nq * A.npad;
  for (int q = 0; q < A.nq; ++q) z[q][r] = zk[(size_t)q * A.npad + k * TB + r];
  __syncthreads();
  const double* T = A.L + ((size_t)m * A.tiles_per_mat + tri_index(i, k)) * TILE_ELEMS;
  double acc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) acc[q] = 0.0;
  for (int c4 = 0; c4 < TB / 4; ++c4) {
    const int s = c4 >> 2, ch = c4 & 3;
    const double2* lp = reinterpret_cast<const double2*>(T + s * SLAB_ELEMS + r * KS + ((ch ^ (r & 3)) << 2));
    const double2 l01 = lp[0], l23 = lp[1];
    const int c = 4 * c4;
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      if (q < A.nq) {
        double v = acc[q];
        v = fma(l01.x, z[q][c], v); v = fma(l01.y, z[q][c + 1], v);
        v = fma(l23.x, z[q][c + 2], v); v = fma(l23.y, z[q][c + 3], v);
        acc[q] = v;
      }
    }
  }
  for (int q = 0; q < A.nq; ++q) rw[(size_t)q * A.npad + i * TB + r] -= acc[q];
}

cudaError_t launch_rl_rhs_update(const CholArgs& a, int k, int nmat, cudaStream_t st, int ntiles) {
  const int gx = k < 0 ? 8 : (ntiles >= 0 ? ntiles : a.nt - 1 - k - a.i0);
  if (gx <= 0) return cudaSuccess;
  rl_rhs_update_kernel<<<dim3(gx, nmat), TB, 0, st>>>(a, k);
  return cudaGetLastError();
}

// ====================================================================== launchers
// $KB200_DIAG_SMEM_KB: shared memory requested per diagonal CTA (KiB, experiments).  Above 113 only ONE diagonal CTA fits an
// SM and its neighbour is a panel CTA of another candidate group (which has the tensor pipe to itself while the diagonal
// CTA is in its strictly sequential pivots) instead of a second diagonal CTA in the same phase.
static size_t diag_smem() {
  static const size_t v = [] {
    const char* e = getenv("KB200_DIAG_SMEM_KB");
    const size_t kb = e ? (size_t)atoi(e) : 0;
    return kb * 1024 > D_SMEM && kb <= 120 ? kb * 1024 : D_SMEM;
  }();
  return v;
}

static cudaError_t half_attrs() {
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(chol_diag_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chol_panel_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(solve_rows_half_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM);
  if (e != cudaSuccess) return e;
  cudaFuncSetAttribute(solve_rows_half_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  e = cudaFuncSetAttribute(solve_rows_half_kernel<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM);
  if (e != cudaSuccess) return e;

  // ask for the largest shared-memory carve-out so that two CTAs fit
  cudaFuncSetAttribute(chol_diag_half_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(chol_panel_half_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done[dev] = true;
  return cudaSuccess;
}

// largest n the in-shared-memory likelihood kernel takes for this input dimension (0: none)
// Measured on B200 against the tile path (tools/time_small.py, 512 candidates, d = 6; profiles/r2k_time_small.log):
// n = 50: 0.054 vs 0.139 ms, n = 128: 0.131 vs 0.150, n = 160: 0.253 vs 0.382, n = 192: 0.329 vs 0.392 -- and
// n = 200, 224 (seven 32-wide panels, one CTA per SM): 0.420 vs 0.401 / 0.406: from N = 224 on the tile path's two
// half-SM CTAs per SM hide the strictly sequential pivots better than one 512-thread CTA can, so it keeps those sizes.
bool half_kernels_skip_padding() { return KB200_RAGGED != 0; }

int lik_small_pad(int n, int d) {
  const int N = (n + 31) & ~31;
  const char* e = getenv("KB200_SMALL_LIK_NMAX");       // (read per call: bench.py measures n = 200 on both paths)
  const int nmax = e ? atoi(e) : 192;
  if (N > nmax || N > 256 || lik_small_smem(N, d) > (size_t)227 * 1024) return 0;
  return N;
}

template <bool FAST, int NTHR>
static cudaError_t lik_small_launch(const CorrArgs& c, const CholArgs& a, int N, size_t sm, int nmat, cudaStream_t st) {
  static size_t attr[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sm > attr[dev]) {
    const int want = (int)(sm > 48 * 1024 ? sm : 48 * 1024);
    cudaError_t e = cudaFuncSetAttribute(lik_small_kernel<FAST, NTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(lik_small_kernel<FAST, NTHR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr[dev] = sm;
  }
  lik_small_kernel<FAST, NTHR><<<nmat, NTHR, sm, st>>>(c, a, N);
  return cudaGetLastError();
}

cudaError_t launch_lik_small(const CorrArgs& c, const CholArgs& a, int nmat, cudaStream_t st) {
  const int N = lik_small_pad(c.n, c.spec.d);
  if (N == 0) return cudaErrorInvalidValue;
  const size_t sm = lik_small_smem(N, c.spec.d);
  const bool fast = c.spec.nkernel == 1 && c.spec.kid[0] == K_GAUSS && !c.spec.kpls;
  const bool wide = N > 128;
  if (fast) return wide ? lik_small_launch<true, 512>(c, a, N, sm, nmat, st) : lik_small_launch<true, 256>(c, a, N, sm, nmat, st);
  return wide ? lik_small_launch<false, 512>(c, a, N, sm, nmat, st) : lik_small_launch<false, 256>(c, a, N, sm, nmat, st);
}

cudaError_t launch_chol_diag_half(const CholArgs& a, int j, int nmat, cudaStream_t st) {
  cudaError_t e = half_attrs();
  if (e != cudaSuccess) return e;
  chol_diag_half_kernel<<<nmat, HT, diag_smem(), st>>>(a, j);
  return cudaGetLastError();
}

// grid_x = number of (half) tiles: 2 (nt - 1 - j) for the factorisation, 2 x point tiles for prediction
cudaError_t launch_chol_panel_half(const CholArgs& a, int j, int grid_x, int grid_y, cudaStream_t st) {
  cudaError_t e = half_attrs();
  if (e != cudaSuccess) return e;
  if (grid_x <= 0 || grid_y <= 0) return cudaSuccess;
  // pure solves (nothing to accumulate before the solve) have a kernel of their own; $KB200_PURE_SOLVE=0: the panel kernel
  static const bool pure_on = !(getenv("KB200_PURE_SOLVE") && atoi(getenv("KB200_PURE_SOLVE")) == 0);
  if (pure_on && (a.rl_mode == 1 || (a.rl_mode == 0 && j == 0))) {
    // CTAs per matrix: about four waves of half-SM CTAs in all, every CTA at least one half tile
    static const int per_launch = getenv("KB200_PURE_SOLVE_CTAS") ? std::max(1, atoi(getenv("KB200_PURE_SOLVE_CTAS"))) : 1184;
    static const int thr = getenv("KB200_PURE_SOLVE_THREADS") ? atoi(getenv("KB200_PURE_SOLVE_THREADS")) : 256;
    const int per = thr > 256 ? per_launch / 2 : per_launch;     // one wide CTA per SM instead of two
    const int nsplit = std::max(1, std::min(grid_x, (per + grid_y - 1) / grid_y));
    if (thr == 640) solve_rows_half_kernel<640><<<dim3(nsplit, grid_y), 640, P_SMEM, st>>>(a, j, grid_x);
    else solve_rows_half_kernel<256><<<dim3(nsplit, grid_y), HT, P_SMEM, st>>>(a, j, grid_x);
    return cudaGetLastError();
  }
  chol_panel_half_kernel<<<dim3(grid_x, grid_y), HT, P_SMEM, st>>>(a, j);
  return cudaGetLastError();
}

}  // namespace kb200
