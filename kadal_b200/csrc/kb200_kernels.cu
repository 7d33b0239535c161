// kadal_b200 — sm_100a kernels for the Kriging hot path.
//
//   corr_tile_kernel     K1  correlation tiles: Psi assembly (likelihood) / psi(X,x*) (predict)
//   chol_diag_kernel     K2  left-looking step j, diagonal tile: DMMA SYRK update, in-smem
//                            POTF2 + inverse, fused forward substitution of [y F], log-det
//   chol_panel_kernel    K3  left-looking step j, tile (i,j): DMMA GEMM update, TRSM as a DMMA
//                            GEMM with the inverted diagonal block.  Prediction points are
//                            "extra rows" of the factorisation and reuse this kernel.
//   lik_finalize_kernel  K4  GLS beta, sigma^2, concentrated / trainvar ln-likelihood
//   trsv_diag / trsv_update  alpha = L^-T r, w = L^-T b (predictor state, once per fit), forward solves
//   append_extract / append_row  one-sample append to a factor (believer / sequential refits, kb200_fit_append)
//   predict_epilogue_kernel  mean / SSqr / s / EI / PoI / PoF / U-function
//   import / export kernels  row-major U, Psi  <->  tile layout
//
// Reference call sites replaced: likelihood_func.py:93-102,168-213; kernelfunc.py:4-94;
// prediction.py:159-171,208-314.  See DESIGN.md for the data layout and rooflines.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "kb200_corr.cuh"
#include "kb200_kernels.h"

namespace kb200 {

// ============================================================================ K1
// One CTA = one 128x128 tile (assembly) or one 128-point row tile x all sample tiles
// (prediction).  256 threads; lane -> (row = lane/4, 32-byte chunk = lane%4), so a warp
// writes 8 consecutive 128-byte lines of a slab (1 KiB contiguous).
constexpr int CORR_THREADS = 256;
// 1: the prediction kernels divide by theta^2 exactly like the likelihood assembly (the first design);
// 0: they multiply by the correctly rounded reciprocal (kb200_corr.cuh:gauss_term)
#ifndef KB200_PREDICT_EXACT_DIV
#define KB200_PREDICT_EXACT_DIV 0
#endif
#ifndef KB200_CORR_FAST_BLOCKS
#define KB200_CORR_FAST_BLOCKS 3
#endif
// 1: the Gaussian fast path computes both rows of a thread per column quad for 8 <= d < 16 (kb200_corr.cuh:corr_quad_gauss2;
// C2 assembly 1.70 -> 1.52 ms per generation, mean prediction at d = 10 9.26 -> 8.26 ms per 2e6 points, same bits); 0: one row at a time
#ifndef KB200_CORR_PAIR2
#define KB200_CORR_PAIR2 1
#endif
#ifndef KB200_CORR_PAIR2_RAGGED   // prediction: the fully real slabs of the ragged last sample tile take the paired path too
#define KB200_CORR_PAIR2_RAGGED 1
#endif
#ifndef KB200_CORR_PAIR2_EDGE   // the same on diagonal tiles / the last tile row of the assembly: measured 1.52 -> 1.82 ms (ptxas spills
#define KB200_CORR_PAIR2_EDGE 0 // 864 instead of 228 bytes at the 85-register cap), off
#endif

__device__ __forceinline__ double standardize1(double x, int normtype, double a, double b) {
  if (normtype == 1) {  // 'default': ((x-lb)/(ub-lb) - 0.5) * 2   samplingplan.py:77-83
    double v = __ddiv_rn(__dsub_rn(x, a), __dsub_rn(b, a));
    return __dmul_rn(__dsub_rn(v, 0.5), 2.0);
  }
  if (normtype == 2) return __ddiv_rn(__dsub_rn(x, a), b);  // (x-mean)/std  prediction.py:164
  return x;
}

// the 2 x 8 quads (1 x 4 entries each) of one thread for one (row tile, column tile) pair.
// INTERIOR: every row and column of the pair is a real sample / point and the pair is not on the
// diagonal, so none of the per-entry fix-ups (identity padding, nugget, zeroed upper triangle, skipped
// quads) can apply: 21 of the 36 tiles of C2, all but the last sample tile of a prediction.
template <int FAST, int MODE, bool INTERIOR>   // FAST: 0 general path; Gaussian fast path: 1..3 = d class 1..3 only (kb200_corr.cuh), 4 = any d
__device__ __forceinline__ void corr_pair(const CorrArgs& A, const double* __restrict__ cand, const double* th2,
                                          double w0, double eps, const double* As, const double* Bt, int lda,
                                          int ti, int tj, double* tile, int warp, int g, int ch,
                                          double (&facc)[2], double (&uacc)[2]) {
  const int d = A.spec.d;
  int s_begin = 0;          // first slab of the one-row-at-a-time loop below
#if KB200_CORR_PAIR2
  if (FAST && (INTERIOR || (KB200_CORR_PAIR2_RAGGED && MODE == CORR_PREDICT))) {
    // both rows of the thread per column quad (corr_quad_gauss2); per entry and per row the same operations in the
    // same order as the loop below.  Prediction against the ragged last sample tile: the slabs whose 16 columns are all real
    // samples take this path too (n = 200: 4 of the 4.5 real slabs of sample tile 1), the rest the loop below.
    const int s_full = INTERIOR ? NSLAB : max(0, min(NSLAB, (A.n - tj * TB) / KS));
    const int row0 = 8 * warp + g, row1 = row0 + 64;
    const double2* thi = reinterpret_cast<const double2*>(th2 + 2 * d);
#pragma unroll 1
    for (int s = 0; s < s_full; ++s) {
      const int c0 = s * KS + 4 * ch;
      Quad q0, q1;
      corr_quad_gauss2<MODE == CORR_ASSEMBLE || KB200_PREDICT_EXACT_DIV, FAST & 3>(d, th2, th2 + d, thi, w0, As + row0 * lda,
                                                                         As + row1 * lda, Bt + c0, TB, q0, q1);
      if (MODE == CORR_ASSEMBLE || tile) {
        double2* dst = reinterpret_cast<double2*>(tile + s * SLAB_ELEMS + row0 * KS + ((ch ^ (row0 & 3)) << 2));
        dst[0] = make_double2(q0.v[0], q0.v[1]);
        dst[1] = make_double2(q0.v[2], q0.v[3]);
        dst = reinterpret_cast<double2*>(tile + s * SLAB_ELEMS + row1 * KS + ((ch ^ (row1 & 3)) << 2));
        dst[0] = make_double2(q1.v[0], q1.v[1]);
        dst[1] = make_double2(q1.v[2], q1.v[3]);
      }
      if (MODE == CORR_PREDICT) {
        const int gc0 = tj * TB + c0;
        double fa0 = 0.0, fa1 = 0.0, ua0 = 0.0, ua1 = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double al = A.alpha[gc0 + e];
          fa0 = fma(q0.v[e], al, fa0);
          fa1 = fma(q1.v[e], al, fa1);
          if (A.wvec) {
            const double wv = A.wvec[gc0 + e];
            ua0 = fma(q0.v[e], wv, ua0);
            ua1 = fma(q1.v[e], wv, ua1);
          }
        }
        facc[0] += fa0; uacc[0] += ua0;
        facc[1] += fa1; uacc[1] += ua1;
      }
    }
    if (INTERIOR) return;
    s_begin = s_full;
  }
  if (KB200_CORR_PAIR2_EDGE && FAST && !INTERIOR && MODE == CORR_ASSEMBLE) {
    // diagonal tiles and the last tile row: the same pairing where both rows of the thread need the quad, then the
    // per-entry fix-ups of the loop below (identity padding, nugget, zeroed upper triangle)
    const int row0 = 8 * warp + g, row1 = row0 + 64;
    const long grow0 = (long)ti * TB + row0, grow1 = grow0 + 64;
    const double2* thi = reinterpret_cast<const double2*>(th2 + 2 * d);
#pragma unroll 1
    for (int s = 0; s < NSLAB; ++s) {
      const int c0 = s * KS + 4 * ch;
      const int gc0 = tj * TB + c0;
      const bool colok = gc0 < A.n;
      const bool n0 = colok && grow0 < A.n && !(ti == tj && c0 > row0);
      const bool n1 = colok && grow1 < A.n && !(ti == tj && c0 > row1);
      Quad q0 = quad_set(0.0), q1 = quad_set(0.0);
      if (n0 && n1) corr_quad_gauss2<true, FAST & 3>(d, th2, th2 + d, thi, w0, As + row0 * lda, As + row1 * lda, Bt + c0, TB, q0, q1);
      else if (n1) q1 = corr_quad_gauss<true, FAST & 3>(d, th2, th2 + d, w0, As + row1 * lda, Bt + c0, TB);
      else if (n0) q0 = corr_quad_gauss<true, FAST & 3>(d, th2, th2 + d, w0, As + row0 * lda, Bt + c0, TB);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gc = gc0 + e;
        if (grow0 >= A.n || gc >= A.n) q0.v[e] = (grow0 == gc) ? 1.0 : 0.0;
        else if (grow0 == gc && A.add_eps) q0.v[e] = __dadd_rn(q0.v[e], eps);
        else if (ti == tj && gc > grow0) q0.v[e] = 0.0;
        if (grow1 >= A.n || gc >= A.n) q1.v[e] = (grow1 == gc) ? 1.0 : 0.0;
        else if (grow1 == gc && A.add_eps) q1.v[e] = __dadd_rn(q1.v[e], eps);
        else if (ti == tj && gc > grow1) q1.v[e] = 0.0;
      }
      double2* dst = reinterpret_cast<double2*>(tile + s * SLAB_ELEMS + row0 * KS + ((ch ^ (row0 & 3)) << 2));
      dst[0] = make_double2(q0.v[0], q0.v[1]);
      dst[1] = make_double2(q0.v[2], q0.v[3]);
      dst = reinterpret_cast<double2*>(tile + s * SLAB_ELEMS + row1 * KS + ((ch ^ (row1 & 3)) << 2));
      dst[0] = make_double2(q1.v[0], q1.v[1]);
      dst[1] = make_double2(q1.v[2], q1.v[3]);
    }
    return;
  }
#endif
#pragma unroll 1
  for (int rgi = 0; rgi < 2; ++rgi) {
    const int row = 8 * (warp + 8 * rgi) + g;
    const long grow = (long)ti * TB + row;
#pragma unroll 1
    for (int s = s_begin; s < NSLAB; ++s) {
      const int c0 = s * KS + 4 * ch;
      const int gc0 = tj * TB + c0;
      Quad q;
      bool need = true;
      if (!INTERIOR) {
        if (MODE == CORR_ASSEMBLE) {
          // skip the strictly-upper part of diagonal tiles and fully padded quads
          if ((ti == tj && c0 > row) || grow >= A.n || gc0 >= A.n) need = false;
        } else {
          if (gc0 >= A.n) need = false;
        }
      }
      if (need) {
        if (FAST) q = corr_quad_gauss<MODE == CORR_ASSEMBLE || KB200_PREDICT_EXACT_DIV, FAST & 3>(d, th2, th2 + d, w0, As + row * lda, Bt + c0, TB);
        else q = corr_quad<MODE == CORR_ASSEMBLE || KB200_PREDICT_EXACT_DIV>(A.spec, cand, As + row * lda, Bt + c0, TB);
      } else {
        q = quad_set(0.0);
      }
      if (!INTERIOR) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int gc = gc0 + e;
          if (MODE == CORR_ASSEMBLE) {
            if (grow >= A.n || gc >= A.n) q.v[e] = (grow == gc) ? 1.0 : 0.0;   // identity padding
            else if (grow == gc && A.add_eps) q.v[e] = __dadd_rn(q.v[e], eps); // Psi + eye*eps
            else if (ti == tj && gc > grow) q.v[e] = 0.0;
          } else {
            if (gc >= A.n) q.v[e] = 0.0;
          }
        }
      }
      // (prediction: the variance kernels never read the columns of the padded 8-column blocks of the last sample tile)
      if (MODE == CORR_ASSEMBLE || (tile && (INTERIOR || !A.skip_pad || gc0 < ((A.n + 7) & ~7)))) {
        double2* dst = reinterpret_cast<double2*>(tile + s * SLAB_ELEMS + row * KS + ((ch ^ (row & 3)) << 2));
        dst[0] = make_double2(q.v[0], q.v[1]);
        dst[1] = make_double2(q.v[2], q.v[3]);
      }
      if (MODE == CORR_PREDICT) {
        double fa = 0.0, ua = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int gc = gc0 + e;   // alpha / w are zero padded to nt*TB
          fa = fma(q.v[e], A.alpha[gc], fa);
          if (A.wvec) ua = fma(q.v[e], A.wvec[gc], ua);
        }
        if (rgi == 0) { facc[0] += fa; uacc[0] += ua; } else { facc[1] += fa; uacc[1] += ua; }
      }
    }
  }
}

template <int FAST, int MODE>
__global__ void __launch_bounds__(CORR_THREADS, FAST ? KB200_CORR_FAST_BLOCKS : 2)
corr_tile_kernel(CorrArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = A.spec.d;
  const int lda = d | 1;
  double* As = reinterpret_cast<double*>(smem_raw);          // [128][lda]
  double* Bt = As + ((TB * lda + 1) & ~1);                   // [d][128]
  double* th2 = Bt + d * TB;                                 // [d] theta^2, [d] RN(1/theta^2), [d] pairs of both
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, ch = lane & 3;

  int ti, tj0, tj1;
  long m;  // candidate (assembly) — prediction uses candidate 0
  if (MODE == CORR_ASSEMBLE) {
    // blockIdx.x -> lower-triangular tile (ti,tj)
    int t = blockIdx.x;
    ti = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    tj0 = t - ti * (ti + 1) / 2;
    tj1 = tj0 + 1;
    m = blockIdx.y;
  } else {
    ti = blockIdx.x;  // point tile
    tj0 = A.tj_split ? blockIdx.y : 0;
    tj1 = A.tj_split ? blockIdx.y + 1 : A.nt;
    m = 0;
  }
  const double* cand = A.cand + m * A.cand_stride;
  if (FAST && (MODE == CORR_ASSEMBLE || KB200_PREDICT_EXACT_DIV) && tid < 2 * d) {   // theta2 | rtheta2 are adjacent; second copy interleaved {theta2_k, rtheta2_k} for one 16-byte load
    const double v = cand[2 * A.spec.ntheta + tid];
    th2[tid] = v;
    th2[2 * d + 2 * (tid % d) + tid / d] = v;
  }
  const double w0 = cand[4 * A.spec.ntheta];

  // ---- row coordinates -> As
  if (MODE == CORR_ASSEMBLE) {
    for (int idx = tid; idx < TB * d; idx += CORR_THREADS) {
      int r = idx / d, k = idx - r * d;
      int gr = ti * TB + r;
      As[r * lda + k] = (gr < A.n) ? A.X[(size_t)gr * d + k] : 0.0;
    }
  } else {
    for (int idx = tid; idx < TB * d; idx += CORR_THREADS) {
      int r = idx / d, k = idx - r * d;
      long gp = (long)ti * TB + r;
      double v = 0.0;
      if (gp < A.npts) {
        v = A.xpts[(size_t)gp * d + k];
        v = standardize1(v, A.normtype, A.norm_a ? A.norm_a[k] : 0.0, A.norm_b ? A.norm_b[k] : 1.0);
        if (FAST && !KB200_PREDICT_EXACT_DIV) v *= A.pscale[k];   // pre-scaled (RN(1/theta_k)): corr_quad_gauss_scaled
      }
      As[r * lda + k] = v;
    }
  }

  double facc[2] = {0.0, 0.0}, uacc[2] = {0.0, 0.0};
  double ftot[2] = {0.0, 0.0}, utot[2] = {0.0, 0.0};
  const double eps = cand[4 * A.spec.ntheta + A.spec.nkernel];

  for (int tj = tj0; tj < tj1; ++tj) {
    __syncthreads();  // As ready / previous Bt consumed
    for (int idx = tid; idx < TB * d; idx += CORR_THREADS) {
      int c = idx / d, k = idx - c * d;
      int gc = tj * TB + c;
      double v = (gc < A.n) ? A.X[(size_t)gc * d + k] : 0.0;
      if (FAST && MODE == CORR_PREDICT && !KB200_PREDICT_EXACT_DIV) v *= A.pscale[k];
      Bt[k * TB + c] = v;
    }
    __syncthreads();
    double* tile;
    bool interior;
    if (MODE == CORR_ASSEMBLE) {
      tile = A.out + ((size_t)m * A.tiles_per_mat + tri_index(ti, tj)) * TILE_ELEMS;
      interior = ti != tj && (ti + 1) * TB <= A.n;
    } else {
      tile = A.out ? A.out + ((size_t)ti * A.nt + tj) * TILE_ELEMS : nullptr;
      interior = (tj + 1) * TB <= A.n;
    }
    if (interior) corr_pair<FAST, MODE, true>(A, cand, th2, w0, eps, As, Bt, lda, ti, tj, tile, warp, g, ch, facc, uacc);
    else corr_pair<FAST, MODE, false>(A, cand, th2, w0, eps, As, Bt, lda, ti, tj, tile, warp, g, ch, facc, uacc);
    if (MODE == CORR_PREDICT) {
      // psi^T alpha, psi^T w per SAMPLE TILE, the tiles added in order: the same bits whether this CTA walks all sample
      // tiles or one CTA per (point tile, sample tile) pair leaves the partial sums to the epilogue (tj_split, small batches)
#pragma unroll
      for (int rgi = 0; rgi < 2; ++rgi) {
        double f = facc[rgi], u = uacc[rgi];
        f += __shfl_xor_sync(0xffffffffu, f, 1);
        f += __shfl_xor_sync(0xffffffffu, f, 2);
        u += __shfl_xor_sync(0xffffffffu, u, 1);
        u += __shfl_xor_sync(0xffffffffu, u, 2);
        ftot[rgi] = tj == tj0 ? f : ftot[rgi] + f;
        utot[rgi] = tj == tj0 ? u : utot[rgi] + u;
        facc[rgi] = 0.0;
        uacc[rgi] = 0.0;
      }
    }
  }
  if (MODE == CORR_PREDICT) {
#pragma unroll
    for (int rgi = 0; rgi < 2; ++rgi) {
      const double f = ftot[rgi], u = utot[rgi];
      const int row = 8 * (warp + 8 * rgi) + g;
      const long gp = (long)ti * TB + row;
      if (ch == 0 && gp < A.npts) {
        const long o = A.tj_split ? (long)blockIdx.y * A.part_stride + gp : gp;
        A.fdot[o] = f;
        if (A.udot) A.udot[o] = u;
      }
    }
  }
}

// ==================================================================== DMMA GEMM core
constexpr int CH_THREADS = 512;     // 16 warps: 4 (M) x 4 (N), warp tile 32x32
constexpr int NSTAGE = 5;           // stage = A slab + B slab = 32 KiB
constexpr int STAGE_ELEMS = 2 * SLAB_ELEMS;
constexpr int LDP = TB + 1;         // POTF2 scratch leading dimension (odd -> conflict free)

struct WarpPos {
  int wm, wn, g, t;
};

__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
}

// one 16-wide k slab (k quarters KQ0 <= kq < KQ1 of it): acc += As(128x16) * Bs(128x16)^T on the
// warp's 32x32 sub-tile
template <int KQ0 = 0, int KQ1 = 4>
__device__ __forceinline__ void mma_slab(double (&acc)[4][4][2], const double* As, const double* Bs,
                                         const WarpPos& w) {
#pragma unroll
  for (int kq = KQ0; kq < KQ1; ++kq) {
    const int sw = ((kq ^ (w.g & 3)) << 2) + w.t;
    double a[4], b[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) a[mt] = As[(32 * w.wm + 8 * mt + w.g) * KS + sw];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[(32 * w.wn + 8 * nt + w.g) * KS + sw];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

// diagonal 32x32 block of a SYRK: only the 8x8 tiles on or below the block diagonal
__device__ __forceinline__ void mma_slab_diag(double (&acc)[4][4][2], const double* As, const WarpPos& w) {
#pragma unroll
  for (int kq = 0; kq < 4; ++kq) {
    const int sw = ((kq ^ (w.g & 3)) << 2) + w.t;
    double a[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) a[mt] = As[(32 * w.wm + 8 * mt + w.g) * KS + sw];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt <= mt; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], a[nt]);
  }
}

// one slab of the SYRK  C(lower) += A A^T  with the 16 warps balanced over the four SM
// sub-partitions (a warp's sub-partition is warp%4 = wm) WITHIN every slab (the slab loop has
// a CTA barrier per slab): the owner (wm > wn) of an off-diagonal block takes k quarters 0-1,
// the otherwise idle mirror warp (wm < wn) takes k quarters 2-3 of the same block (split-K,
// summed by syrk_reduce), and diagonal warps skip the tiles above the diagonal.  Work per
// sub-partition and slab: 136 DMMA instead of 256.
__device__ __forceinline__ void syrk_slab(double (&acc)[4][4][2], const double* As, const WarpPos& w) {
  if (w.wm > w.wn) {
    mma_slab<0, 2>(acc, As, As, w);
  } else if (w.wm < w.wn) {
    const WarpPos h{w.wn, w.wm, w.g, w.t};
    mma_slab<2, 4>(acc, As, As, h);
  } else {
    mma_slab_diag(acc, As, w);
  }
}

// helpers hand their partial block to the owners through shared memory (red: 6 x 1024 doubles)
__device__ __forceinline__ void syrk_reduce(double (&acc)[4][4][2], double* red, const WarpPos& w, int lane) {
  __syncthreads();
  const int hi = w.wm > w.wn ? w.wm : w.wn, lo = w.wm > w.wn ? w.wn : w.wm;
  double* buf = red + (hi * (hi - 1) / 2 + lo) * 1024 + lane;
  if (w.wm < w.wn) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        buf[((mt * 4 + nt) * 2 + 0) * 32] = acc[mt][nt][0];
        buf[((mt * 4 + nt) * 2 + 1) * 32] = acc[mt][nt][1];
      }
  }
  __syncthreads();
  if (w.wm > w.wn) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[mt][nt][0] += buf[((mt * 4 + nt) * 2 + 0) * 32];
        acc[mt][nt][1] += buf[((mt * 4 + nt) * 2 + 1) * 32];
      }
  }
  __syncthreads();
}

// acc += A(128 x 16*nslab) * B(128 x 16*nslab)^T (SYRK = false) or the lower blocks of A A^T
// (SYRK = true, one operand), streamed slab by slab from HBM/L2 with cp.async.bulk into an
// NSTAGE-deep ring.  full[st] (tx-count mbarrier) says "slab landed"; empty[st] (16 arrivals,
// one per warp) says "every warp is done with it", so there is NO CTA-wide barrier per slab:
// warps drift freely and the producer (thread 0) refills a stage two slabs after it was
// consumed, when its empty barrier has normally completed already.  `it` is the running slab
// counter (identical in all threads) that defines stage and parity; `stage_elems` is the stage
// stride in doubles.  Callers must __syncthreads() before reusing the stage memory.
constexpr int LOOKAHEAD = NSTAGE - 2;
template <bool SYRK, typename OnSlab>
__device__ __forceinline__ void gemm_stream(double (&acc)[4][4][2], const double* __restrict__ gA,
                                            const double* __restrict__ gB, int nslab,
                                            double* stages, int stage_elems, uint64_t* full, uint32_t& it,
                                            const WarpPos& w, OnSlab on_slab) {
  const int tid = threadIdx.x;
  uint64_t* empty = full + NSTAGE + 2;
  auto issue = [&](int q, uint32_t cur) {
    const int st = cur % NSTAGE;
    const uint32_t use = cur / NSTAGE;
    if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
    double* dst = stages + (size_t)st * stage_elems;
    mbar_expect_tx(&full[st], SYRK ? SLAB_BYTES : 2 * SLAB_BYTES);
    bulk_g2s(dst, gA + (size_t)q * SLAB_ELEMS, SLAB_BYTES, &full[st]);
    if (!SYRK) bulk_g2s(dst + SLAB_ELEMS, gB + (size_t)q * SLAB_ELEMS, SLAB_BYTES, &full[st]);
  };
  if (tid == 0) {
    const int pro = nslab < LOOKAHEAD ? nslab : LOOKAHEAD;
    for (int q = 0; q < pro; ++q) issue(q, it + q);
  }
  for (int q = 0; q < nslab; ++q) {
    const uint32_t cur = it + q;
    const int st = cur % NSTAGE;
    if (tid == 0 && q + LOOKAHEAD < nslab) issue(q + LOOKAHEAD, cur + LOOKAHEAD);
    mbar_wait(&full[st], (cur / NSTAGE) & 1);
    const double* As = stages + (size_t)st * stage_elems;
    if (SYRK) syrk_slab(acc, As, w);
    else mma_slab(acc, As, As + SLAB_ELEMS, w);
    on_slab(As, q);
    // release one slab late: every DMMA / FMA of slab q - 1 has been issued by now, so its operand loads
    // have returned and the bulk copy that re-fills the stage cannot land under them (kb200_half.cu)
    if (q > 0) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[(cur - 1) % NSTAGE]);
    }
    if (SYRK && (q & 7) == 7) __syncthreads();   // tile boundary: publishes the staged Z tile
  }
  if (nslab > 0) {     // hand back the last slab too (a later call continues on the same barriers)
    fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[(it + nslab - 1) % NSTAGE]);
  }
  it += nslab;
}

struct NoSlabOp {
  __device__ __forceinline__ void operator()(const double*, int) const {}
};
// pulls a 128 KiB tile into L2 a few slabs before the GEMM phase ends (2 lines per thread): the
// C tile is only needed afterwards, and at kernel start would be evicted again by the ~2.4 GB/ms
// the 148 SMs stream through the 126 MB L2 (profile r1h)
struct PrefetchTileOp {
  const char* p;
  int at;
  __device__ __forceinline__ void operator()(const double*, int q) const {
    if (q == at) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
    }
  }
};

// barrier block: full[NSTAGE] + wfull[2] (one arrival + tx bytes), then empty[NSTAGE] (one
// arrival per warp)
constexpr int NBARRIERS = 2 * NSTAGE + 2;
__device__ __forceinline__ void init_barriers(uint64_t* bars) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE + 2; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&bars[NSTAGE + 2 + i], CH_THREADS / 32);
    fence_mbar_init();
  }
  __syncthreads();
}

// naive (debug) GEMM: same contract as gemm_stream, plain global loads + DFMA
__device__ __forceinline__ void gemm_naive(double (&acc)[4][4][2], const double* gA, const double* gB, int nslab,
                           const WarpPos& w) {
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 32 * w.wm + 8 * mt + w.g, c = 32 * w.wn + 8 * nt + 2 * w.t + e;
        double s = 0.0;
        for (int q = 0; q < nslab; ++q)
          for (int k = 0; k < KS; ++k)
            s = fma(gA[(size_t)q * SLAB_ELEMS + slab_off(r, k)], gB[(size_t)q * SLAB_ELEMS + slab_off(c, k)], s);
        acc[mt][nt][e] += s;
      }
}

// ------------------------------------------------- in-smem blocked POTF2 (diagonal tile)
// The 128x128 diagonal tile is factorised inside one CTA with a blocked right-looking
// algorithm (block 32):
//   (1) warp 0 factorises the 32x32 diagonal block in REGISTERS (lane = row, rank-1 updates
//       broadcast with warp shuffles) and inverts its four 8x8 diagonal sub-blocks (W8);
//   (2) the rows below (and the right-hand-side strip [y F]^T, which rides along as eight extra
//       rows) are solved against the block 8 rows per warp with DMMA: a right-looking TRSM
//       whose only "divisions" are products with the 8x8 inverses;
//   (3) the trailing part is updated with DMMA (SYRK, lower 16x16 items + RHS items).
// No 128x128 inverse is ever formed: the panel kernel runs the same right-looking TRSM with
// the sixteen 8x8 inverses of the tile.
//
// S (LDP-strided) holds the lower triangle of the SPD tile on entry and L on exit; the part
// above the diagonal is scratch.  trow (TLD-strided, 8 rows) holds the RHS strip.
constexpr int TLD = TB + 4;
constexpr unsigned FULLMASK = 0xffffffffu;

// C-fragment (c0,c1 = T[g][2t], T[g][2t+1]) of an 8x8 tile  ->  the two A-fragments
// (a0 = T[g][t], a1 = T[g][4+t]) of the same tile, g = lane/4, t = lane%4.
__device__ __forceinline__ void cfrag_to_afrag(double c0, double c1, double& a0, double& a1, int lane) {
  const int src0 = (lane & ~3) + ((lane & 3) >> 1);
  const double v00 = __shfl_sync(FULLMASK, c0, src0), v01 = __shfl_sync(FULLMASK, c1, src0);
  const double v10 = __shfl_sync(FULLMASK, c0, src0 + 2), v11 = __shfl_sync(FULLMASK, c1, src0 + 2);
  a0 = (lane & 1) ? v01 : v00;
  a1 = (lane & 1) ? v11 : v10;
}

// sqrt(x) and 1/sqrt(x) of a positive normal double: MUFU.RSQ64H seed (>= 20 bits) and two
// coupled Newton (Goldschmidt) steps, s refined once more -> both within ~1 ulp.  This is the
// latency-critical operation of the factorisation (one per pivot, strictly sequential).
__device__ __forceinline__ void sqrt_rsqrt(double x, double& s, double& twice_h) {
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
  twice_h = h;            // caller multiplies by 2 (folded into its own operand)
}

// warp-level Cholesky of the 32x32 block at B (stride LDP): lane r owns row r in registers.
// Column k of L is broadcast through shared memory (lbuf, 2 x 32 doubles, double buffered: one
// STS + wide LDS per lane instead of 31 double shuffles: 2x faster, tools/microbench).  A lone
// warp retires roughly one FP64 instruction per 8 cycles here, so the step cost follows the
// instruction count (variants with a shorter dependency chain but more arithmetic were slower).
// dinv[k] = 1/L_kk.  Returns first bad pivot (global index + 1) or 0.
__device__ __noinline__ int potf2_32_warp(double* B, double* dinv, double* lbuf, int lane, int base) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? B[lane * LDP + c] : 0.0;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    double piv = __shfl_sync(FULLMASK, a[k], k);
    const bool ok = piv > 0.0;
    bad = (!ok && bad == 0) ? base + k + 1 : bad;
    piv = ok ? piv : 1.0;
    const double a2 = a[k] + a[k];
    double dk, h;
    sqrt_rsqrt(piv, dk, h);
    const double l = a2 * h;                           // a[k] / sqrt(piv)
    double* lb = lbuf + (k & 1) * 32;
    lb[lane] = l;
    __syncwarp();
#pragma unroll
    for (int c = k + 1; c < 32; ++c) a[c] = fma(-l, lb[c], a[c]);
    a[k] = (lane == k) ? dk : l;
    if (lane == k) dinv[k] = h + h;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) B[lane * LDP + c] = a[c];
  return bad;
}

// inverses of the four 8x8 diagonal sub-blocks of the 32x32 block at B (one warp:
// lane = (sub-block, column)).  out: 4 x 64 doubles, row-major 8x8 each, zeros above.
__device__ __forceinline__ void inv8_warp(const double* B, const double* dinv, double* out, int lane) {
  const int blk = lane >> 3, c = lane & 7;
  const double* Lb = B + (8 * blk) * LDP + 8 * blk;
  double w[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < r; ++t) s = fma(Lb[r * LDP + t], w[t], s);
    w[r] = ((r == c) ? 1.0 : -s) * dinv[8 * blk + r];
    if (r < c) w[r] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) out[blk * 64 + r * 8 + c] = w[r];
}

// one 8-row strip against the 32x32 block: rows <- rows * L_kk^-T (right-looking, DMMA).
// rows: pointer to element (first row of the strip, first column of the block), stride ld.
__device__ __forceinline__ void strip_trsm32(double* rows, int ld, const double* Lkk, const double* W8,
                                             int lane) {
  const int g = lane >> 2, tq = lane & 3;
  double x[4][2];
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    x[ct][0] = rows[g * ld + 8 * ct + 2 * tq];
    x[ct][1] = rows[g * ld + 8 * ct + 2 * tq + 1];
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
        const double b0 = Lkk[(8 * ct + g) * LDP + 8 * t + tq];
        const double b1 = Lkk[(8 * ct + g) * LDP + 8 * t + 4 + tq];
        dmma884(x[ct][0], x[ct][1], a0, b0);
        dmma884(x[ct][0], x[ct][1], a1, b1);
      }
    }
  }
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    rows[g * ld + 8 * ct + 2 * tq] = x[ct][0];
    rows[g * ld + 8 * ct + 2 * tq + 1] = x[ct][1];
  }
}

// trailing update after the 32-wide panel at column o (items it0 <= it < it1 of the list
// "lower 16x16 blocks in row-major triangular order, then the RHS 8x16 blocks"), done by
// `nw` warps of which this one is number `widx`:
//   S[r][c]   -= sum_k S[r][o+k] * S[c][o+k]        o+32 <= c <= r      (16x16 items, lower)
//   trow[q][c] -= sum_k trow[q][o+k] * S[c][o+k]                         (8x16 RHS items)
__device__ __forceinline__ void trailing_update32(double* S, double* trow, int o, int it0, int it1,
                                                  int widx, int nw, int lane) {
  const int g = lane >> 2, tq = lane & 3;
  const int base = o + 32;
  const int nb = (TB - base) >> 4;
  const int ntri = nb * (nb + 1) / 2;
  const int nitems = ntri + nb;
  if (it1 > nitems) it1 = nitems;
  for (int it = it0 + widx; it < it1; it += nw) {
    if (it < ntri) {
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= it) ++bi;
      const int bj = it - bi * (bi + 1) / 2;
      double* ar0 = S + (base + 16 * bi + g) * LDP;
      double* ar1 = ar0 + 8 * LDP;
      const double* br0 = S + (base + 16 * bj + g) * LDP;
      const double* br1 = br0 + 8 * LDP;
      const int cc = base + 16 * bj + 2 * tq;
      double c[2][2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        c[0][nt][0] = ar0[cc + 8 * nt]; c[0][nt][1] = ar0[cc + 8 * nt + 1];
        c[1][nt][0] = ar1[cc + 8 * nt]; c[1][nt][1] = ar1[cc + 8 * nt + 1];
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
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (bi > bj || nt == 0) { ar0[cc + 8 * nt] = c[0][nt][0]; ar0[cc + 8 * nt + 1] = c[0][nt][1]; }
        ar1[cc + 8 * nt] = c[1][nt][0]; ar1[cc + 8 * nt + 1] = c[1][nt][1];
      }
    } else {
      const int bj = it - ntri;
      double* ar0 = trow + g * TLD;
      const double* br0 = S + (base + 16 * bj + g) * LDP;
      const double* br1 = br0 + 8 * LDP;
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

// Blocked factorisation of the tile in S with the RHS strip in trow.  On exit: S lower = L,
// dinv[128] = 1/L_kk, W8s = sixteen 8x8 inverse diagonal blocks, trow = (L^-1 [y F])^T.
// s_bad (shared) receives the first non-positive pivot index + 1.  All threads must call.
// Look-ahead: of the trailing update of panel kb only the three items that touch the next
// diagonal block are done before warp 0 starts on it; the rest runs on warps 1..15 while
// warp 0 is inside the (strictly sequential, latency-bound) 32x32 factorisation.
__device__ void potf2_blocked(double* S, double* trow, double* dinv, double* W8s, double* lbuf, int* s_bad) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid == 0) *s_bad = 0;
  __syncthreads();
  for (int kb = 0; kb < 4; ++kb) {
    const int o = 32 * kb;
    if (warp == 0) {
      const int b = potf2_32_warp(S + o * LDP + o, dinv + o, lbuf, lane, o);
      __syncwarp();
      inv8_warp(S + o * LDP + o, dinv + o, W8s + kb * 256, lane);
      if (lane == 0 && b && *s_bad == 0) *s_bad = b;
    } else if (kb > 0) {
      trailing_update32(S, trow, o - 32, 3, 1 << 20, warp - 1, nwarp - 1, lane);
    }
    __syncthreads();
    const int nrow_strips = (TB - o - 32) >> 3;
    for (int st = warp; st <= nrow_strips; st += nwarp) {
      if (st < nrow_strips) strip_trsm32(S + (o + 32 + 8 * st) * LDP + o, LDP, S + o * LDP + o, W8s + kb * 256, lane);
      else strip_trsm32(trow + o, TLD, S + o * LDP + o, W8s + kb * 256, lane);
    }
    __syncthreads();
    if (kb < 3) {
      trailing_update32(S, trow, o, 0, 3, warp, nwarp, lane);
      __syncthreads();
    }
  }
}

// plain column-by-column variant (debug / KB200_FLAG_NAIVE): same outputs, DFMA only
__device__ void potf2_simple(double* S, double* trow, double* dinv, double* W8s, int* s_bad, int nq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarp = nthreads >> 5;
  int bad = 0;
  for (int k = 0; k < TB; ++k) {
    __syncthreads();
    double piv = S[k * LDP + k];
    if (!(piv > 0.0)) {
      if (!bad) bad = k + 1;
      piv = 1.0;
    }
    const double dk = sqrt(piv), inv = 1.0 / dk;
    __syncthreads();
    for (int r = k + 1 + tid; r < TB; r += nthreads) S[r * LDP + k] *= inv;
    if (tid == 0) { S[k * LDP + k] = dk; dinv[k] = inv; }
    __syncthreads();
    for (int c = k + 1 + warp; c < TB; c += nwarp) {
      const double lc = S[c * LDP + k];
      for (int r = c + lane; r < TB; r += 32) S[r * LDP + c] = fma(-S[r * LDP + k], lc, S[r * LDP + c]);
    }
  }
  __syncthreads();
  if (tid == 0) *s_bad = bad;
  if (tid < 128) {
    const int blk = tid >> 3, c = tid & 7;
    const double* Lb = S + (8 * blk) * LDP + 8 * blk;
    double w[8];
    for (int r = 0; r < 8; ++r) {
      double s = 0.0;
      for (int t = 0; t < r; ++t) s = fma(Lb[r * LDP + t], w[t], s);
      w[r] = (r < c) ? 0.0 : ((r == c) ? 1.0 : -s) * dinv[8 * blk + r];
    }
    for (int r = 0; r < 8; ++r) W8s[blk * 64 + r * 8 + c] = w[r];
  }
  if (tid >= 128 && tid < 128 + nq) {
    double* t = trow + (tid - 128) * TLD;
    for (int r = 0; r < TB; ++r) {
      double s = t[r];
      for (int c = 0; c < r; ++c) s = fma(-S[r * LDP + c], t[c], s);
      t[r] = s * dinv[r];
    }
  }
  __syncthreads();
}

// write L (lower incl. diagonal, zeros above) from the scratch to tile layout
__device__ void store_L(const double* S, double* Ltile, int nthreads) {
  for (int idx = threadIdx.x; idx < TILE_ELEMS / 4; idx += nthreads) {
    // idx -> (slab, row, chunk)
    const int ch = idx & 3, row = (idx >> 2) & (TB - 1), s = idx >> 9;
    const int c0 = s * KS + 4 * ch;
    double l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int c = c0 + e;
      l[e] = (c <= row) ? S[row * LDP + c] : 0.0;
    }
    const int off = s * SLAB_ELEMS + row * KS + ((ch ^ (row & 3)) << 2);
    double2* dl = reinterpret_cast<double2*>(Ltile + off);
    dl[0] = make_double2(l[0], l[1]);
    dl[1] = make_double2(l[2], l[3]);
  }
}

// ============================================================================ K2
// Diagonal-tile work shared by chol_diag_kernel and the panel kernel's "next diagonal" role.
// Shared-memory map: [0, 132 KiB) S scratch | ... | trow 8xTLD | dinv 128 | W8 out 1024.
struct DiagSmem {
  double* S;       // [128][LDP] POTF2 scratch
  double* trow;    // [8][TLD]  RHS strip
  double* dinv;    // [128]
  double* W8s;     // [16][64]
  double* lbuf;    // [2][32] column broadcast buffer of potf2_32_warp
  int* s_bad;
  double* s_red;   // [16]
};

// fused forward substitution hook: t[r] -= L_jk[r][c] * Z_k[c]   (likelihood_func.py:183-193)
// thread -> (row rr = tid/4, column quarter hh = tid%4) of the slab
struct RhsHook {
  double (&tacc)[MAXQ];
  const double* zb;     // Z of this matrix, [nq][npad]
  int npad, nq, rr, hh, col0;
  __device__ __forceinline__ void operator()(const double* As, int q) const {
    const double2* lp = reinterpret_cast<const double2*>(As + rr * KS + ((hh ^ (rr & 3)) << 2));
    const double2 l01 = lp[0], l23 = lp[1];
    const int col = col0 + q * KS + 4 * hh;
#pragma unroll
    for (int qq = 0; qq < MAXQ; ++qq) {
      if (qq < nq) {
        const double* z = zb + (size_t)qq * npad + col;
        double s = tacc[qq];
        s = fma(l01.x, z[0], s); s = fma(l01.y, z[1], s);
        s = fma(l23.x, z[2], s); s = fma(l23.y, z[3], s);
        tacc[qq] = s;
      }
    }
  }
};

// The same hook with Z staged in shared memory: a rolling two-tile buffer zs[2][nq][128].  The
// per-slab global loads of the plain hook sit in front of the slab barrier with their full
// L2 latency exposed (44 % of all stall samples of the panel kernel's diagonal role, profile
// r1e); here tile t+1 is fetched into registers at the first slab of tile t and stored to the
// other half of the buffer at its last slab (the slab barrier publishes it).
struct RhsHookStaged {
  double (&tacc)[MAXQ];
  double (&zreg)[2];
  const double* zg;     // Z of this matrix, [nq][npad] (global)
  double* zs;           // shared, [2][nq][128]
  int npad, nq, rr, hh, tile0, ntile;
  __device__ __forceinline__ void preload(int tile) const {
    for (int idx = threadIdx.x; idx < nq * TB; idx += CH_THREADS)
      zs[((tile & 1) * nq) * TB + idx] = zg[(size_t)(idx >> 7) * npad + tile * TB + (idx & (TB - 1))];
  }
  __device__ __forceinline__ void operator()(const double* As, int q) const {
    const int tile = tile0 + (q >> 3), sq = q & 7;
    const bool more = tile + 1 < ntile;
    if (sq == 0 && more) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = threadIdx.x + e * CH_THREADS;
        if (idx < nq * TB) zreg[e] = zg[(size_t)(idx >> 7) * npad + (tile + 1) * TB + (idx & (TB - 1))];
      }
    }
    const double2* lp = reinterpret_cast<const double2*>(As + rr * KS + ((hh ^ (rr & 3)) << 2));
    const double2 l01 = lp[0], l23 = lp[1];
    const double* zt = zs + ((tile & 1) * nq) * TB + sq * KS + 4 * hh;
#pragma unroll
    for (int qq = 0; qq < MAXQ; ++qq) {
      if (qq < nq) {
        const double2 z01 = *reinterpret_cast<const double2*>(zt + qq * TB);
        const double2 z23 = *reinterpret_cast<const double2*>(zt + qq * TB + 2);
        double s = tacc[qq];
        s = fma(l01.x, z01.x, s); s = fma(l01.y, z01.y, s);
        s = fma(l23.x, z23.x, s); s = fma(l23.y, z23.y, s);
        tacc[qq] = s;
      }
    }
    if (sq == 7 && more) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = threadIdx.x + e * CH_THREADS;
        if (idx < nq * TB) zs[(((tile + 1) & 1) * nq) * TB + idx] = zreg[e];
      }
    }
  }
};

// From the SYRK accumulators to the finished diagonal tile jd of matrix m:
//   S = Psi_jd,jd - acc; t = R_jd - tacc; blocked POTF2; outputs L_jd,jd (in place), the 8x8
//   inverse blocks, Z_jd and the log-det partial.  All threads of the CTA must call.
template <bool NAIVE>
__device__ void diag_finish(const CholArgs& A, long m, int jd, double (&acc)[4][4][2], double (&tacc)[MAXQ],
                            const DiagSmem& D, const WarpPos& w) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = A.nq, rr = tid >> 2, hh = tid & 3;
  double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
  double* Cjj = tiles + (size_t)tri_index(jd, jd) * TILE_ELEMS;
  double* S = D.S;
  double* trow = D.trow;
  for (int idx = tid; idx < 8 * TLD; idx += CH_THREADS) trow[idx] = 0.0;
  __syncthreads();
  if (w.wm >= w.wn) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = 32 * w.wm + 8 * mt + w.g, c = 32 * w.wn + 8 * nt + 2 * w.t;
        const double2 p = *reinterpret_cast<const double2*>(Cjj + tile_off(r, c));
        if (c <= r) S[r * LDP + c] = p.x - acc[mt][nt][0];
        if (c + 1 <= r) S[r * LDP + c + 1] = p.y - acc[mt][nt][1];
      }
  }
  // RHS strip: t = R_jd - sum_k L_jd,k Z_k   (row q of trow)
#pragma unroll
  for (int qq = 0; qq < MAXQ; ++qq) {
    if (qq < nq) {
      double s = tacc[qq];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (hh == 0) trow[qq * TLD + rr] = A.R[(size_t)qq * A.npad + jd * TB + rr] - s;
    }
  }
  __syncthreads();
  if (NAIVE) potf2_simple(S, trow, D.dinv, D.W8s, D.s_bad, nq);
  else potf2_blocked(S, trow, D.dinv, D.W8s, D.lbuf, D.s_bad);
  const int bad = *D.s_bad;
  if (tid == 0 && bad && A.info[m] == 0) A.info[m] = jd * TB + bad;
  // ---- outputs: L_jj, the 8x8 inverse blocks, Z_j, log-det partial
  store_L(S, Cjj, CH_THREADS);
  {
    double* w8 = A.W8 + ((size_t)m * A.nt + jd) * 1024;
    for (int idx = tid; idx < 1024; idx += CH_THREADS) w8[idx] = D.W8s[idx];
  }
  for (int idx = tid; idx < nq * TB; idx += CH_THREADS) {
    const int q = idx >> 7, r = idx & (TB - 1);
    A.Z[((size_t)m * A.nq + q) * A.npad + jd * TB + r] = trow[q * TLD + r];
  }
  {
    double v = (tid < TB) ? log(fabs(S[tid * LDP + tid])) : 0.0;   // padded rows are exactly 1
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) D.s_red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < 4; ++i) s += D.s_red[i];
      A.logdet[m] = (jd == 0 ? 0.0 : A.logdet[m]) + s;
    }
  }
}

constexpr int DIAG_EXTRA_ELEMS = 8 * TLD + TB + 1024 + 16 + 64;

template <bool NAIVE>
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_diag_kernel(CholArgs A, int j) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);                 // NSTAGE*32 KiB, aliased by S
  DiagSmem D;
  D.S = stages;
  D.trow = stages + (size_t)NSTAGE * STAGE_ELEMS;                       // [8][TLD]
  D.dinv = D.trow + 8 * TLD;                                            // [128]
  D.W8s = D.dinv + TB;                                                  // [16][64]
  D.s_red = D.W8s + 1024;                                               // [16]
  D.lbuf = D.s_red + 16;                                                // [64]
  uint64_t* full = reinterpret_cast<uint64_t*>(D.lbuf + 64);
  D.s_bad = reinterpret_cast<int*>(full + NBARRIERS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WarpPos w{warp & 3, warp >> 2, lane >> 2, lane & 3};
  const long m = blockIdx.x;
  const double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
  const double* rowj = tiles + (size_t)tri_index(j, 0) * TILE_ELEMS;
  const int nq = A.nq;
  const double* zb = A.Z + (size_t)m * A.nq * A.npad;
  init_barriers(full);

  double acc[4][4][2];
  zero_acc(acc);
  double tacc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) tacc[q] = 0.0;
  const int rr = tid >> 2, hh = tid & 3;

  uint32_t it = 0;
  const int nslab = NSLAB * j;
  if (NAIVE) {
    gemm_naive(acc, rowj, rowj, nslab, w);
    for (int q = 0; q < nslab; ++q)
      for (int e = 0; e < 4; ++e) {
        double lv = rowj[(size_t)q * SLAB_ELEMS + slab_off(rr, 4 * hh + e)];
#pragma unroll
        for (int qq = 0; qq < MAXQ; ++qq)
          if (qq < nq) tacc[qq] = fma(lv, zb[(size_t)qq * A.npad + q * KS + 4 * hh + e], tacc[qq]);
      }
  } else if (nslab > 0) {
    gemm_stream<true>(acc, rowj, rowj, nslab, stages, STAGE_ELEMS, full, it, w,
                      RhsHook{tacc, zb, A.npad, nq, rr, hh, 0});
    syrk_reduce(acc, stages, w, lane);
  }
  diag_finish<NAIVE>(A, m, j, acc, tacc, D, w);
}

// 8x8 inverse blocks of an imported factor (predictor_load path): one CTA per tile column
__global__ void __launch_bounds__(128)
w8_from_tiles_kernel(CholArgs A) {
  const int j = blockIdx.x;
  const long m = blockIdx.y;
  const double* Ljj = A.L + ((size_t)m * A.tiles_per_mat + tri_index(j, j)) * TILE_ELEMS;
  const int blk = threadIdx.x >> 3, c = threadIdx.x & 7;
  double w[8];
  for (int r = 0; r < 8; ++r) {
    double s = 0.0;
    for (int t = 0; t < r; ++t) s = fma(Ljj[tile_off(8 * blk + r, 8 * blk + t)], w[t], s);
    const double d = Ljj[tile_off(8 * blk + r, 8 * blk + r)];
    w[r] = (r < c) ? 0.0 : ((r == c) ? 1.0 : -s) / (d != 0.0 ? d : 1.0);
  }
  double* w8 = A.W8 + ((size_t)m * A.nt + j) * 1024;
  for (int r = 0; r < 8; ++r) w8[blk * 64 + r * 8 + c] = w[r];
}

// ============================================================================ K3
// tile (i,j) of the factor (MODE_CHOL) or tile (pt,j) of a prediction chunk (MODE_PRED):
//   X = (C - A_row * B_row^T) * L_jj^-T
// GEMM update: 4x4 warps of 32x32 (DMMA, operands streamed by bulk copies).  TRSM: the tile is
// re-distributed through shared memory so that each warp owns 8 complete rows (rows are
// independent), then solved right-looking column tile by column tile: the finished 8x8 tile
// times the 8x8 inverse block, followed by DMMA updates of the column tiles to its right.
//
// "Next diagonal" role (A.fuse_diag, CTA with i == j+1): once L_{j+1,j} is known the whole
// tile row j+1 is final, so the same CTA goes on to the diagonal tile j+1: SYRK over the row
// (tiles 0..j-1 re-streamed from L2, tile j straight from the shared-memory copy of X), then
// diag_finish (POTF2, 8x8 inverses, Z_{j+1}, log-det).  No separate diagonal launches.
//
// Shared memory (doubles):  [0, 5*4096)  GEMM stages (2 slabs each); Cs/Xs = first 16384,
//   Ws = stage 4;  SYRK stages (1 slab each) at [16384, 16384 + 5*2048);  W8s at 26624;
//   diag extras (trow, dinv, s_red) behind it; S scratch aliases [0, 128*LDP).
constexpr int PANEL_W8_OFF = TILE_ELEMS + NSTAGE * SLAB_ELEMS;          // 26624 doubles = 208 KiB
constexpr int PANEL_EXTRA_OFF = PANEL_W8_OFF + 1024;
static_assert(PANEL_W8_OFF >= NSTAGE * STAGE_ELEMS, "W8s must not overlap the GEMM stages");
static_assert(TB * LDP <= PANEL_W8_OFF, "POTF2 scratch must end before W8s");
static_assert(2 * MAXQ * TB <= 1024 + 8 * TLD, "staged Z tiles must fit in the W8s + trow area");

template <bool NAIVE>
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_panel_kernel(CholArgs A, int j) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);          // 5 x 32 KiB
  double* Cs = stages;                                           // 128 KiB (aliases stages 0..3)
  double* Ws = stages + 4 * STAGE_ELEMS;                         // 2 x 16 KiB (aliases stage 4)
  double* W8s = stages + PANEL_W8_OFF;                           // [16][64]
  DiagSmem D;
  D.S = stages;
  D.trow = stages + PANEL_EXTRA_OFF;
  D.dinv = D.trow + 8 * TLD;
  D.W8s = W8s;
  D.s_red = D.dinv + TB;
  D.lbuf = D.s_red + 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(D.lbuf + 64);
  uint64_t* wfull = full + NSTAGE;
  D.s_bad = reinterpret_cast<int*>(full + NBARRIERS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WarpPos w{warp & 3, warp >> 2, lane >> 2, lane & 3};
  const double *rowA, *rowB, *Ljj, *W8j;
  double* Cij;
  long ssq_base = 0;
  long m = 0;
  bool next_diag = false;
  if (A.pred_mode) {
    const long pt = blockIdx.x;
    double* base = A.chunk + (size_t)pt * A.nt * TILE_ELEMS;
    rowA = base;
    Cij = base + (size_t)j * TILE_ELEMS;
    rowB = A.L + (size_t)tri_index(j, 0) * TILE_ELEMS;
    Ljj = A.L + (size_t)tri_index(j, j) * TILE_ELEMS;
    W8j = A.W8 + (size_t)j * 1024;
    ssq_base = pt * TB;
  } else {
    const int i = j + 1 + blockIdx.x;
    m = blockIdx.y;
    double* tiles = A.L + (size_t)m * A.tiles_per_mat * TILE_ELEMS;
    rowA = tiles + (size_t)tri_index(i, 0) * TILE_ELEMS;
    rowB = tiles + (size_t)tri_index(j, 0) * TILE_ELEMS;
    Cij = tiles + (size_t)tri_index(i, j) * TILE_ELEMS;
    Ljj = tiles + (size_t)tri_index(j, j) * TILE_ELEMS;
    W8j = A.W8 + ((size_t)m * A.nt + j) * 1024;
    next_diag = A.fuse_diag && blockIdx.x == 0;
  }
  init_barriers(full);

  double acc[4][4][2];
  zero_acc(acc);
  uint32_t it = 0;
  if (NAIVE) gemm_naive(acc, rowA, rowB, NSLAB * j, w);
  else gemm_stream<false>(acc, rowA, rowB, NSLAB * j, stages, STAGE_ELEMS, full, it, w,
                          PrefetchTileOp{reinterpret_cast<const char*>(Cij) + (size_t)tid * 256,
                                         NSLAB * j > 6 ? NSLAB * j - 6 : 0});

  // ---- C = Psi_ij - acc  -> Cs (tile layout)
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r = 32 * w.wm + 8 * mt + w.g, c = 32 * w.wn + 8 * nt + 2 * w.t;
      const int off = tile_off(r, c);
      const double2 p = *reinterpret_cast<const double2*>(Cij + off);
      *reinterpret_cast<double2*>(Cs + off) = make_double2(p.x - acc[mt][nt][0], p.y - acc[mt][nt][1]);
    }
  // (Ws / W8s are refilled by bulk copies after generic-proxy reads: proxy fence + CTA barrier)
  fence_proxy_async();
  __syncthreads();

  if (NAIVE) {
    // one thread per row: x[c] = (C[r][c] - sum_{t<c} x[t] L[c][t]) / L[c][c]
    if (tid < TB) {
      for (int c = 0; c < TB; ++c) {
        double s = Cs[tile_off(tid, c)];
        for (int t = 0; t < c; ++t) s = fma(-Cs[tile_off(tid, t)], Ljj[tile_off(c, t)], s);
        Cs[tile_off(tid, c)] = s / Ljj[tile_off(c, c)];
      }
    }
    __syncthreads();
  } else {
    if (tid == 0) {
      mbar_expect_tx(&wfull[0], SLAB_BYTES + 8192);
      bulk_g2s(W8s, W8j, 8192, &wfull[0]);
      bulk_g2s(Ws, Ljj, SLAB_BYTES, &wfull[0]);
      mbar_expect_tx(&wfull[1], SLAB_BYTES);
      bulk_g2s(Ws + SLAB_ELEMS, Ljj + SLAB_ELEMS, SLAB_BYTES, &wfull[1]);
    }
  }
  // ---- every warp takes 8 complete rows of the tile (C-fragment layout, 16 column tiles)
  const int g = w.g, tq = w.t;
  const int xr = 8 * warp + g;
  double x[16][2];
#pragma unroll
  for (int ct = 0; ct < 16; ++ct) {
    const double2 p = *reinterpret_cast<const double2*>(Cs + tile_off(xr, 8 * ct + 2 * tq));
    x[ct][0] = p.x;
    x[ct][1] = p.y;
  }
  if (!NAIVE) {
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) {
      const int stg = s & 1;
      mbar_wait(&wfull[stg], (s >> 1) & 1);
      const double* Ls = Ws + (size_t)stg * SLAB_ELEMS;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int t = 2 * s + tt;
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
            const double* lrow = Ls + (8 * ct + g) * KS + tq;
            const double b0 = lrow[((2 * tt) ^ (g & 3)) << 2];
            const double b1 = lrow[((2 * tt + 1) ^ (g & 3)) << 2];
            dmma884(x[ct][0], x[ct][1], a0, b0);
            dmma884(x[ct][0], x[ct][1], a1, b1);
          }
        }
      }
      fence_proxy_async();   // generic-proxy reads of this slot before the bulk copy that re-fills it
      __syncthreads();
      if (tid == 0 && s + 2 < NSLAB) {
        mbar_expect_tx(&wfull[stg], SLAB_BYTES);
        bulk_g2s(Ws + (size_t)stg * SLAB_ELEMS, Ljj + (size_t)(s + 2) * SLAB_ELEMS, SLAB_BYTES, &wfull[stg]);
      }
    }
  }
  // ---- store X (tile layout) and, for prediction, accumulate ||x_row||^2
  double rs = 0.0;
#pragma unroll
  for (int ct = 0; ct < 16; ++ct) {
    const double2 xv = make_double2(x[ct][0], x[ct][1]);
    const int off = tile_off(xr, 8 * ct + 2 * tq);
    *reinterpret_cast<double2*>(Cij + off) = xv;
    if (next_diag) *reinterpret_cast<double2*>(Cs + off) = xv;     // Xs: operand of the SYRK below
    rs = fma(x[ct][0], x[ct][0], rs);
    rs = fma(x[ct][1], x[ct][1], rs);
  }
  if (A.pred_mode) {
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    if (tq == 0) {
      double* dst = A.ssq + ssq_base + xr;
      *dst = (j == 0 ? 0.0 : *dst) + rs;
    }
  }
  if (!next_diag) return;

  // ================= next diagonal tile jd = j + 1 of matrix m (CTA-uniform branch)
  const int jd = j + 1;
  const double* zb = A.Z + (size_t)m * A.nq * A.npad;
  zero_acc(acc);
  double tacc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) tacc[q] = 0.0;
  const int rr = tid >> 2, hh = tid & 3;
  __syncthreads();       // Xs complete
  if (NAIVE) {
    // not used (the naive path keeps separate diagonal launches)
  } else {
    // Z tiles are staged in the (now free) W8s/trow area; tile 0 first
    double zreg[2] = {0.0, 0.0};
    double* zs = W8s;
    const RhsHookStaged hook0{tacc, zreg, zb, zs, A.npad, A.nq, rr, hh, 0, jd};
    hook0.preload(0);
    __syncthreads();
    // row jd, tiles 0..j-1: streamed; one-slab stages behind Xs
    gemm_stream<true>(acc, rowA, rowA, NSLAB * j, stages + TILE_ELEMS, SLAB_ELEMS, full, it, w, hook0);
    // tile j: the X just computed, read in place from shared memory
    const RhsHookStaged hook{tacc, zreg, zb, zs, A.npad, A.nq, rr, hh, j, jd};
#pragma unroll 1
    for (int q = 0; q < NSLAB; ++q) {
      const double* As = Cs + (size_t)q * SLAB_ELEMS;
      syrk_slab(acc, As, w);
      hook(As, q);
    }
    syrk_reduce(acc, stages + TILE_ELEMS, w, lane);
  }
  diag_finish<NAIVE>(A, m, jd, acc, tacc, D, w);
}

// ============================================================================ K4
// beta = (B^T B)^-1 B^T a, r = a - B beta, q = r^T r, with a = L^-1 y, B = L^-1 F
// (Z rows).  Equivalent to likelihood_func.py:180-204 (<=1e-13 rel).
__device__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(256)
lik_finalize_kernel(FinArgs A) {
  __shared__ double red[8];
  __shared__ double G[(MAXQ - 1) * (MAXQ - 1)], gv[MAXQ - 1], beta[MAXQ - 1];
  const long m = blockIdx.x;
  const int p = A.nq - 1, npad = A.npad;
  const double* a = A.Z + (size_t)m * A.nq * npad;
  const double* B = a + npad;
  // G = B^T B, g = B^T a
  for (int i = 0; i < p; ++i) {
    for (int k = 0; k <= i; ++k) {
      double v = 0.0;
      for (int r = threadIdx.x; r < npad; r += blockDim.x) v = fma(B[(size_t)i * npad + r], B[(size_t)k * npad + r], v);
      v = block_sum(v, red);
      if (threadIdx.x == 0) G[i * p + k] = G[k * p + i] = v;
    }
    double v = 0.0;
    for (int r = threadIdx.x; r < npad; r += blockDim.x) v = fma(B[(size_t)i * npad + r], a[r], v);
    v = block_sum(v, red);
    if (threadIdx.x == 0) gv[i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // p x p SPD solve (Cholesky), p <= 7
    double Lg[(MAXQ - 1) * (MAXQ - 1)];
    for (int i = 0; i < p; ++i)
      for (int k = 0; k <= i; ++k) {
        double s = G[i * p + k];
        for (int t = 0; t < k; ++t) s -= Lg[i * p + t] * Lg[k * p + t];
        Lg[i * p + k] = (i == k) ? sqrt(s) : s / Lg[k * p + k];
      }
    double yv[MAXQ - 1];
    for (int i = 0; i < p; ++i) {
      double s = gv[i];
      for (int t = 0; t < i; ++t) s -= Lg[i * p + t] * yv[t];
      yv[i] = s / Lg[i * p + i];
    }
    for (int i = p - 1; i >= 0; --i) {
      double s = yv[i];
      for (int t = i + 1; t < p; ++t) s -= Lg[t * p + i] * beta[t];
      beta[i] = s / Lg[i * p + i];
    }
  }
  __syncthreads();
  double v = 0.0;
  for (int r = threadIdx.x; r < npad; r += blockDim.x) {
    double res = a[r];
    for (int i = 0; i < p; ++i) res = fma(-B[(size_t)i * npad + r], beta[i], res);
    if (A.resid) A.resid[(size_t)m * npad + r] = res;
    v = fma(res, res, v);
  }
  const double q = block_sum(v, red);
  if (threadIdx.x == 0) {
    const double* cand = A.cand + m * A.cand_stride;
    const double sig_tv = cand[A.cand_stride - 1];
    const double half_logdet = A.logdet[m];   // sum ln L_ii = 0.5 * LnDetPsi
    double nll, sig2;
    if (A.trainvar) {
      sig2 = sig_tv;
      nll = half_logdet + 0.5 * A.nsamp * log(2.0 * 3.141592653589793) + 0.5 * A.nsamp * log(sig2) + q / (2.0 * sig2);
    } else {
      sig2 = q / (double)A.n;
      nll = 0.5 * (double)A.n * log(sig2) + half_logdet;
    }
    if (A.info[m] != 0 || !(nll == nll)) nll = INFINITY;
    A.nll[m] = nll;
    if (A.sigma2) A.sigma2[m] = sig2;
    if (A.beta) for (int i = 0; i < p; ++i) A.beta[m * p + i] = beta[i];
    if (A.btb) A.btb[m] = G[0];
  }
}

// ================================================================== backsolve (fit)
// X_j = L_jj^-T (RHS_j - sum_{k>j} L_kj^T X_k), j = nt-1 .. 0 (and the forward counterpart), nq columns.
// (once per fit: alpha = L^-T r, w = L^-T b of the predictor, prediction.py:233-235, 245-249)
// The in-tile solve is a column-oriented substitution in shared memory: thread (r, q-group).
// The diagonal tile is staged in shared memory first (it used to be read from global memory inside the 128
// strictly sequential steps, ~2 us per step: 8 of the 10 ms of a fit at n = 4000) and the pivots' reciprocals
// are formed once, off the chain.
constexpr size_t SOLVE_SMEM = (size_t)(TILE_ELEMS + 1024 + 8 * MAXQ) * 8;
// Blocked by the 8x8 diagonal blocks whose inverses W8 the factorisation keeps anyway: x_b = W_b t_b (or W_b^T t_b),
// then the block is taken out of the rows still to be solved -- 16 block steps of two barriers per tile instead of
// 128 pivots of one (41 -> 15 us per tile).
__device__ __forceinline__ void tile_substitute(const double* __restrict__ T, const double* __restrict__ W8, double* Ts,
                                                double* tv, double* out_base, int npad, int nq, bool transposed) {
  // forward (transposed = false): solve L x = t;  backward (true): solve L^T x = t
  const int tid = threadIdx.x, r = tid & (TB - 1), part = tid >> 7;
  double* Ws = Ts + TILE_ELEMS;          // [16][8][8]
  double* xb = Ws + 1024;                // [8][MAXQ]
  {
    const double2* src = reinterpret_cast<const double2*>(T);
    double2* dst = reinterpret_cast<double2*>(Ts);
    for (int idx = tid; idx < TILE_ELEMS / 2; idx += CH_THREADS) dst[idx] = src[idx];
    for (int idx = tid; idx < 1024; idx += CH_THREADS) Ws[idx] = W8[idx];
  }
  __syncthreads();
  for (int step = 0; step < TB / 8; ++step) {
    const int b = transposed ? TB / 8 - 1 - step : step;
    if (r < 8) {
      for (int q = part; q < nq; q += 4) {
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double w = transposed ? Ws[b * 64 + k * 8 + r] : Ws[b * 64 + r * 8 + k];
          x = fma(w, tv[(8 * b + k) * MAXQ + q], x);
        }
        xb[r * MAXQ + q] = x;
        out_base[(size_t)q * npad + 8 * b + r] = x;
      }
    }
    __syncthreads();
    const bool mine = transposed ? (r < 8 * b) : (r >= 8 * b + 8);
    if (mine) {
      double l[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) l[k] = transposed ? Ts[tile_off(8 * b + k, r)] : Ts[tile_off(r, 8 * b + k)];
      for (int q = part; q < nq; q += 4) {
        double t = tv[r * MAXQ + q];
#pragma unroll
        for (int k = 0; k < 8; ++k) t = fma(-l[k], xb[k * MAXQ + q], t);
        tv[r * MAXQ + q] = t;
      }
    }
    __syncthreads();
  }
}

// Right-looking over tile columns, two launches per column: trsv_diag_kernel solves the diagonal tile for the nq
// columns (one CTA, 128 sequential pivots in shared memory), trsv_update_kernel takes that block out of every
// remaining right-hand-side block (one CTA per tile, all in parallel).  The first version was ONE CTA walking the
// whole triangle (6 ms of the 8 ms of a fit at n = 4000, ncu launch list r1r).
__global__ void __launch_bounds__(CH_THREADS, 1)
trsv_diag_kernel(BackArgs A, int j, int transposed) {
  extern __shared__ __align__(16) double solve_tile[];   // diagonal tile + reciprocal pivots (SOLVE_SMEM)
  __shared__ double tv[TB * MAXQ];
  const int nq = A.nq, npad = A.npad;
  for (int idx = threadIdx.x; idx < TB * nq; idx += CH_THREADS) {
    const int c = idx / nq, q = idx - c * nq;
    tv[c * MAXQ + q] = A.work[(size_t)q * npad + j * TB + c];
  }
  __syncthreads();
  tile_substitute(A.L + (size_t)tri_index(j, j) * TILE_ELEMS, A.W8 + (size_t)j * 1024, solve_tile, tv, A.out + j * TB, npad,
                  nq, transposed != 0);
}

// forward:  work_i -= L_ij   x_j, i = j + 1 + blockIdx.x   (tile (i, j), rows of the tile)
// backward: work_i -= L_ji^T x_j, i = blockIdx.x < j       (tile (j, i), columns of the tile)
__global__ void __launch_bounds__(CH_THREADS)
trsv_update_kernel(BackArgs A, int j, int transposed) {
  __shared__ double xs[TB * MAXQ];
  __shared__ double part[4][TB][MAXQ];
  const int tid = threadIdx.x, r = tid & (TB - 1), part_id = tid >> 7;
  const int nq = A.nq, npad = A.npad;
  const int i = transposed ? (int)blockIdx.x : j + 1 + (int)blockIdx.x;
  for (int idx = tid; idx < TB * nq; idx += CH_THREADS) {
    const int c = idx / nq, q = idx - c * nq;
    xs[c * MAXQ + q] = A.out[(size_t)q * npad + j * TB + c];
  }
  __syncthreads();
  const double* T = A.L + (size_t)(transposed ? tri_index(j, i) : tri_index(i, j)) * TILE_ELEMS;
  double acc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) acc[q] = 0.0;
  if (transposed) {
    // out index r = column of the tile: row c of the tile is contiguous over r
#pragma unroll 8
    for (int c = 32 * part_id; c < 32 * part_id + 32; ++c) {
      const double l = T[tile_off(c, r)];
#pragma unroll
      for (int q = 0; q < MAXQ; ++q)
        if (q < nq) acc[q] = fma(l, xs[c * MAXQ + q], acc[q]);
    }
  } else {
    // out index r = row of the tile: this thread reads the two 128-byte lines (slabs 2 part, 2 part + 1) of its row
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int slab = 2 * part_id + sl;
      const double2* line = reinterpret_cast<const double2*>(T + (size_t)slab * SLAB_ELEMS + r * KS);
      double v[KS];
#pragma unroll
      for (int h = 0; h < KS / 2; ++h) {
        const double2 t = line[h];
        v[2 * h] = t.x;
        v[2 * h + 1] = t.y;
      }
#pragma unroll
      for (int pp = 0; pp < KS; ++pp) {
        const int c = slab * KS + ((((pp >> 2) ^ (r & 3)) << 2) | (pp & 3));   // inverse of slab_off's swizzle
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
          if (q < nq) acc[q] = fma(v[pp], xs[c * MAXQ + q], acc[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MAXQ; ++q)
    if (q < nq) part[part_id][r][q] = acc[q];
  __syncthreads();
  if (tid < TB) {
    for (int q = 0; q < nq; ++q) {
      const double sum = ((part[0][r][q] + part[1][r][q]) + part[2][r][q]) + part[3][r][q];
      A.work[(size_t)q * npad + i * TB + r] -= sum;
    }
  }
}

// ========================================================================= decode
// x -> per-candidate parameter rows.  Mirrors likelihood_func.py:57-79 and nuggetset
// (:121-165): layout [log10 theta | log10 sigma2 (trainvar) | log10 nugget (tuned) |
// kernel weights (nkernel>1)] decided by nbhyp.
__global__ void decode_kernel(DecodeArgs A) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= A.P) return;
  const double* x = A.x + m * A.nbhyp;
  double* c = A.cand + m * A.cand_stride;
  const int nt = A.ntheta, nk = A.nkernel;
  for (int k = 0; k < nt; ++k) {
    const double xk = A.iso ? x[0] : x[k];
    const double th = A.theta_is_linear ? xk : pow(10.0, xk);
    const double th2 = __dmul_rn(th, th);
    c[k] = th;
    c[nt + k] = __drcp_rn(th);
    c[2 * nt + k] = th2;
    c[3 * nt + k] = __drcp_rn(th2);
  }
  const int off = (A.iso ? 1 : nt) + (A.trainvar ? 1 : 0);
  double eps = A.eps_fixed;
  if (A.layout & 1) eps = pow(10.0, x[off]);             // tuned nugget
  const int woff = off + ((A.layout & 1) ? 1 : 0);
  if (A.layout & 2) {                                    // composite weights / sum
    double s = 0.0;
    for (int k = 0; k < nk; ++k) s = __dadd_rn(s, x[woff + k]);
    for (int k = 0; k < nk; ++k) c[4 * nt + k] = __ddiv_rn(x[woff + k], s);
  } else if (A.wgkf_given) {
    for (int k = 0; k < nk; ++k) c[4 * nt + k] = A.wgkf_given[k];
  } else {
    for (int k = 0; k < nk; ++k) c[4 * nt + k] = 1.0;
  }
  c[4 * nt + nk] = eps;
  c[4 * nt + nk + 1] = A.trainvar ? pow(10.0, x[A.iso ? 1 : nt]) : 0.0;
}

// ============================================================ import / export (fit)
__global__ void import_U_kernel(const double* __restrict__ U, int n, int nt, double* __restrict__ tiles) {
  // tile layout L from row-major upper U (= L^T); identity padding
  const int t = blockIdx.x;
  int ti = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  double* T = tiles + (size_t)t * TILE_ELEMS;
  for (int idx = threadIdx.x; idx < TILE_ELEMS; idx += blockDim.x) {
    int c = idx >> 7, r = idx & (TB - 1);   // r fastest: U[gc][gr] contiguous in gr
    int gr = ti * TB + r, gc = tj * TB + c;
    double v;
    if (gr < n && gc < n) v = (gc <= gr) ? U[(size_t)gc * n + gr] : 0.0;
    else v = (gr == gc) ? 1.0 : 0.0;
    T[tile_off(r, c)] = v;
  }
}

__global__ void export_kernel(const double* __restrict__ tiles, int n, double* __restrict__ out, int what) {
  // what = 0: U row-major upper (U[r][c] = L[c][r], zeros below the diagonal)
  // what = 1: full symmetric Psi
  const size_t total = (size_t)n * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
    double v;
    if (what == 0) {
      v = (c >= r) ? tiles[(size_t)tri_index(c / TB, r / TB) * TILE_ELEMS + tile_off(c % TB, r % TB)] : 0.0;
    } else {
      int hi = r >= c ? r : c, lo = r >= c ? c : r;
      v = tiles[(size_t)tri_index(hi / TB, lo / TB) * TILE_ELEMS + tile_off(hi % TB, lo % TB)];
    }
    out[idx] = v;
  }
}

__global__ void add_diag_kernel(double* tiles, int n, const double* eps_ptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double eps = *eps_ptr;
    double* p = tiles + (size_t)tri_index(i / TB, i / TB) * TILE_ELEMS + tile_off(i % TB, i % TB);
    *p = __dadd_rn(*p, eps);
  }
}

// ======================================================================= epilogue
// prediction.py:233-235 (mean), :245-251 (SSqr, s), :258-310 (acquisition values),
// akmcs.py:281-285 (U-function)
__device__ __forceinline__ void epilogue_point(const EpiArgs& A, long i, double fdot, double udot, double ssq,
                                               long& zero_ss, int& any) {
  const double tiny = 2.2250738585072014e-308;
  const double RSQ2 = 0.7071067811865475;        // 1/np.sqrt(2)
  const double SQ2PI = 2.5066282746310002;       // np.sqrt(2*np.pi)
  const double f = A.beta + fdot;
  if (A.pred) A.pred[i] = f;
  if (!A.need_var) return;
  const double u = udot - 1.0;
  const double term1 = 1.0 - ssq;
  const double term2 = u * (u / A.btb);
  const double ssqr = A.sigma2 * (term1 + term2);
  const double s = sqrt(fabs(ssqr));
  if (A.ssqr) A.ssqr[i] = ssqr;
  if (A.s) A.s[i] = s;
  if (ssqr == 0.0) ++zero_ss;
  if (A.ei) {
    const double dlt = A.ymin - f;
    const double t1 = dlt * (0.5 + 0.5 * erf(RSQ2 * dlt / s));
    const double t2 = s / SQ2PI * exp(-0.5 * (dlt * dlt) / ssqr);
    if (t1 != 0.0 || t2 != 0.0) any = 1;
    A.ei[i] = -(t1 + t2 + tiny);
  }
  if (A.poi) A.poi[i] = -(0.5 + 0.5 * erf(RSQ2 * (A.ymin - f) / s));
  if (A.pof) {
    const double z = A.limit_ge ? (f - A.limit) / s : (A.limit - f) / s;
    A.pof[i] = 0.5 + 0.5 * erf(RSQ2 * z);
  }
  if (A.ufun) A.ufun[i] = fabs(f) / s;
}

__global__ void predict_epilogue_kernel(EpiArgs A) {
  long zero_ss = 0;
  int any = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < A.npts; i += (long)gridDim.x * blockDim.x) {
    double f = A.fdot[i], u = A.need_var ? A.udot[i] : 0.0;
    for (int p = 1; p < A.nparts; ++p) {           // per-sample-tile partial sums, added in tile order (deterministic)
      f += A.fdot[(long)p * A.part_stride + i];
      if (A.need_var) u += A.udot[(long)p * A.part_stride + i];
    }
    epilogue_point(A, i, f, u, A.need_var ? A.ssq[i] : 0.0, zero_ss, any);
  }
  if (A.stats) {
    if (zero_ss) atomicAdd(reinterpret_cast<unsigned long long*>(A.stats), (unsigned long long)zero_ss);
    if (any) atomicOr(reinterpret_cast<unsigned long long*>(A.stats + 1), 1ull);
  }
}

// ================================================================ fused small-n predictor
// n <= 256 (one or two sample tiles), one Gaussian kernel, d <= 8: the whole of
// prediction.py:159-251 for 128 points per CTA pass in ONE kernel -- standardise, psi in
// DMMA C-fragment layout straight into registers (each warp owns 8 points), psi^T alpha and
// psi^T w, right-looking DMMA TRSM against L_00, DMMA GEMM against L_10 with V_0 as the A
// operand from shared memory, TRSM against L_11, ||v||^2 and the epilogue.  Nothing of size
// n x npred touches HBM (the general path writes the psi chunk and re-reads it per tile
// column: 29 GB per 4e6 points at n = 200).  The factor slabs stream from L2 through a 3-stage
// cp.async.bulk ring shared by the 16 warps; the schedule runs ahead across point tiles.
constexpr int PS_VLD = TB + 4;
constexpr int PS_NST = 3;
constexpr int PS_LOOK = 1;    // with the one-slab-late release below the ring needs NST >= LOOK + 2
constexpr int PS_MAXD = 8;

// sum_k (a_k - b_k)^2 / theta_k^2 for two neighbouring samples, numpy's order for d <= 8
__device__ __forceinline__ void gauss_pair_sum(int d, const double* th2, const double* a, const double* bt2,
                                               double& S0, double& S1) {
  if (!KB200_PREDICT_EXACT_DIV) {   // coordinates pre-scaled by RN(1/theta_k) (kb200_corr.cuh:corr_quad_gauss_scaled)
    // a plain loop over d: the per-d unrolled switch this replaced was slower (code size, see psi_tile)
    S0 = 0.0; S1 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double ak = a[k];
      const double2 b = *reinterpret_cast<const double2*>(bt2 + k * 256);
      const double d0 = ak - b.x, d1 = ak - b.y;
      S0 = fma(d0, d0, S0);
      S1 = fma(d1, d1, S1);
    }
    return;
  }
  auto term = [&](int k, double& t0, double& t1) {
    const double ak = a[k];
    const double2 b = *reinterpret_cast<const double2*>(bt2 + k * 256);
    const double t2 = th2[k], r2 = th2[PS_MAXD + k];
    const double d0 = __dsub_rn(ak, b.x), d1 = __dsub_rn(ak, b.y);
    if (KB200_PREDICT_EXACT_DIV) {
      t0 = div_by_rcp(__dmul_rn(d0, d0), t2, r2);
      t1 = div_by_rcp(__dmul_rn(d1, d1), t2, r2);
    } else {            // prediction: d^2 * RN(1/theta^2), see gauss_term
      t0 = __dmul_rn(__dmul_rn(d0, d0), r2);
      t1 = __dmul_rn(__dmul_rn(d1, d1), r2);
    }
  };
  if (d < 8) {
    S0 = 0.0; S1 = 0.0;
    for (int k = 0; k < d; ++k) {
      double t0, t1;
      term(k, t0, t1);
      S0 = __dadd_rn(S0, t0);
      S1 = __dadd_rn(S1, t1);
    }
  } else {   // d == 8: ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
    double p[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) term(k, p[k], q[k]);
    S0 = __dadd_rn(__dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3])), __dadd_rn(__dadd_rn(p[4], p[5]), __dadd_rn(p[6], p[7])));
    S1 = __dadd_rn(__dadd_rn(__dadd_rn(q[0], q[1]), __dadd_rn(q[2], q[3])), __dadd_rn(__dadd_rn(q[4], q[5]), __dadd_rn(q[6], q[7])));
  }
}

// right-looking TRSM of an 8-row strip (C fragments x) against a diagonal tile of which only the
// first `nslab` slabs matter; its slabs arrive as ring entries seq0 + s.  The column bound NCT
// is compile time on purpose: with a run-time column bound x[] drops out of the register file
// (272 bytes of stack, 85 M local-memory wavefronts in profile r1i).
template <int NCT, typename Acquire, typename Release>
__device__ __forceinline__ void trsm_strip(double (&x)[16][2], int seq0, int nslab, const double* W8, int lane,
                                           Acquire acquire, Release release) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int s = 0; s < NCT / 2; ++s) {
    if (s >= nslab) break;               // trailing slabs hold identity padding only
    const double* Ls = acquire(seq0 + s);
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
      const int t = 2 * s + tt;
      double a0, a1;
      cfrag_to_afrag(x[t][0], x[t][1], a0, a1, lane);
      double y0 = 0.0, y1 = 0.0;
      dmma884(y0, y1, a0, W8[t * 64 + g * 8 + tq]);
      dmma884(y0, y1, a1, W8[t * 64 + g * 8 + 4 + tq]);
      x[t][0] = y0;
      x[t][1] = y1;
      if (t < NCT - 1) {
        cfrag_to_afrag(y0, y1, a0, a1, lane);
        a0 = -a0;
        a1 = -a1;
#pragma unroll
        for (int ct = t + 1; ct < NCT; ++ct) {
          const double* lrow = Ls + (8 * ct + g) * KS + tq;
          dmma884(x[ct][0], x[ct][1], a0, lrow[((2 * tt) ^ (g & 3)) << 2]);
          dmma884(x[ct][0], x[ct][1], a1, lrow[((2 * tt + 1) ^ (g & 3)) << 2]);
        }
      }
    }
    release(seq0 + s);
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1)
predict_small_kernel(PredSmallArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = A.d, lda = d | 1, n = A.n, nt = A.nt;
  double* Vs = reinterpret_cast<double*>(smem_raw);              // [128][PS_VLD]  V_0 = L_00^-1 psi_0
  double* stages = Vs + TB * PS_VLD;                             // PS_NST slabs
  double* W8s = stages + PS_NST * SLAB_ELEMS;                    // [2][1024]
  double* Bt = W8s + 2048;                                       // [d][256] samples, transposed, zero padded
  double* alph = Bt + PS_MAXD * 256;                             // [256]
  double* wv = alph + 256;                                       // [256]
  double* th2 = wv + 256;                                        // [PS_MAXD] theta^2 | [PS_MAXD] 1/theta^2
  double* As = th2 + 2 * PS_MAXD;                                // [128][lda] standardised points
  uint64_t* full = reinterpret_cast<uint64_t*>(As + TB * (PS_MAXD + 1));
  uint64_t* empty = full + PS_NST;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
  const int row = 8 * warp + g;
  const int ntiles = (int)((A.npts + TB - 1) / TB);
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nct1 = nt == 2 ? (n - TB + 7) >> 3 : 0;              // valid 8-column tiles of sample tile 1
  const int ns1 = (nct1 + 1) >> 1;                               // slabs of L_11 that matter
  const int NS = nt == 2 ? 16 + ns1 : 8;                         // slabs per point tile
  const int total = my_tiles * NS;

  if (tid == 0) {
    for (int i = 0; i < PS_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CH_THREADS / 32); }
    fence_mbar_init();
  }
  for (int idx = tid; idx < nt * 1024; idx += CH_THREADS) W8s[idx] = A.W8[idx];
  for (int idx = tid; idx < d * 256; idx += CH_THREADS) {
    const int k = idx >> 8, c = idx & 255;
    double v = c < n ? A.X[(size_t)c * d + k] : 0.0;
    if (!KB200_PREDICT_EXACT_DIV) v *= A.pscale[k];
    Bt[idx] = v;
  }
  for (int idx = tid; idx < 256; idx += CH_THREADS) {
    alph[idx] = idx < nt * TB ? A.alpha[idx] : 0.0;
    wv[idx] = idx < nt * TB ? A.wvec[idx] : 0.0;
  }
  if (KB200_PREDICT_EXACT_DIV && tid < d) { th2[tid] = A.cand[2 * A.ntheta + tid]; th2[PS_MAXD + tid] = A.cand[3 * A.ntheta + tid]; }
  const double w0 = A.cand[4 * A.ntheta];
  __syncthreads();

  auto issue = [&](int seq) {          // thread 0 only
    const int i = seq % NS, st = seq % PS_NST;
    const int use = seq / PS_NST;
    if (use > 0) mbar_wait(&empty[st], (uint32_t)((use - 1) & 1));
    const double* src = i < 8 ? A.L + (size_t)i * SLAB_ELEMS
                      : i < 16 ? A.L + (size_t)TILE_ELEMS + (size_t)(i - 8) * SLAB_ELEMS
                               : A.L + (size_t)2 * TILE_ELEMS + (size_t)(i - 16) * SLAB_ELEMS;
    mbar_expect_tx(&full[st], SLAB_BYTES);
    bulk_g2s(stages + (size_t)st * SLAB_ELEMS, src, SLAB_BYTES, &full[st]);
  };
  auto acquire = [&](int seq) -> const double* {
    if (tid == 0 && seq + PS_LOOK < total) issue(seq + PS_LOOK);
    const int st = seq % PS_NST;
    mbar_wait(&full[st], (uint32_t)((seq / PS_NST) & 1));
    return stages + (size_t)st * SLAB_ELEMS;
  };
  // release one slab late (entry seq - 1 when the work on entry seq has been issued): its DMMAs precede
  // this point, so their operand loads have returned before the bulk copy can re-fill the stage
  auto release = [&](int seq) {
    if (seq > 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[(seq - 1) % PS_NST]);
    }
  };
  if (tid == 0)
    for (int q = 0; q < PS_LOOK && q < total; ++q) issue(q);

  // psi of sample tile `tile` for this thread's row: columns 8 ct + 2 tq (+1), ct < nct
  double x[16][2];
  double fa, ua;
  // (The fused predictor's loop body is large enough for instruction fetch to show up in the stall samples -- ncu r2g:
  // no_instruction 5-7 % -- and smaller code is faster here: pair sums as a loop over d instead of a per-d unrolled switch
  // 3.42e8 -> 3.60e8 pts/s at C3.  Rolling this column-tile loop as well, with a switch steering the results into x[],
  // shrinks the kernel from 12.6 k to 8.0 k instructions but makes ptxas spill x[]: 2.40e8.)
  auto psi_tile = [&](int tile, int nct) {
    const double* arow = As + row * lda;
#pragma unroll
    for (int ct = 0; ct < 16; ++ct) {
      double e0 = 0.0, e1 = 0.0;
      if (ct < nct) {
        const int col = TB * tile + 8 * ct + 2 * tq;
        double S0, S1;
        gauss_pair_sum(d, th2, arow, Bt + col, S0, S1);
        if (KB200_PREDICT_EXACT_DIV) {
          e0 = col < n ? __dmul_rn(w0, exp(-0.5 * S0)) : 0.0;
          e1 = col + 1 < n ? __dmul_rn(w0, exp(-0.5 * S1)) : 0.0;
        } else {
          e0 = col < n ? w0 * exp_nonpos(-0.5 * S0) : 0.0;
          e1 = col + 1 < n ? w0 * exp_nonpos(-0.5 * S1) : 0.0;
        }
        const double2 al = *reinterpret_cast<const double2*>(alph + col);
        const double2 ww = *reinterpret_cast<const double2*>(wv + col);
        fa = fma(e0, al.x, fa); fa = fma(e1, al.y, fa);
        ua = fma(e0, ww.x, ua); ua = fma(e1, ww.y, ua);
      }
      x[ct][0] = e0;
      x[ct][1] = e1;
    }
  };
  long zero_ss = 0;
  int any = 0;
  int iter = 0;
  for (int pt = blockIdx.x; pt < ntiles; pt += gridDim.x, ++iter) {
    __syncthreads();                      // every warp is done with the previous tile's As
    for (int idx = tid; idx < TB * d; idx += CH_THREADS) {
      const int r = idx / d, k = idx - r * d;
      const long gp = (long)pt * TB + r;
      double v = 0.0;
      if (gp < A.npts) v = standardize1(A.xpts[(size_t)gp * d + k], A.normtype, A.norm_a ? A.norm_a[k] : 0.0, A.norm_b ? A.norm_b[k] : 1.0);
      if (!KB200_PREDICT_EXACT_DIV) v *= A.pscale[k];
      As[r * lda + k] = v;
    }
    __syncthreads();
    const int seq = iter * NS;
    fa = 0.0; ua = 0.0;
    double ssq = 0.0;
    // one pass per sample tile; every helper is expanded exactly once so that x[] stays in
    // registers (a lambda called from two places is outlined and x[] would live on the stack)
#pragma unroll 1
    for (int tile = 0; tile < nt; ++tile) {
      const int nct = tile == 0 ? 16 : nct1;
      psi_tile(tile, nct);
      if (tile == 1) {
        // x = psi_1 - V_0 L_10^T : A fragments of V_0 straight from shared memory
        const double* vrow = Vs + row * PS_VLD + tq;
#pragma unroll 1
        for (int s = 0; s < NSLAB; ++s) {
          const double* Ls = acquire(seq + 8 + s);
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const int t = 2 * s + tt;
            const double a0 = -vrow[8 * t], a1 = -vrow[8 * t + 4];
#pragma unroll
            for (int ct = 0; ct < 16; ++ct) {
              if (ct < nct1) {
                const double* lrow = Ls + (8 * ct + g) * KS + tq;
                dmma884(x[ct][0], x[ct][1], a0, lrow[((2 * tt) ^ (g & 3)) << 2]);
                dmma884(x[ct][0], x[ct][1], a1, lrow[((2 * tt + 1) ^ (g & 3)) << 2]);
              }
            }
          }
          release(seq + 8 + s);
        }
      }
      // (one call site: with several instantiations behind branches ptxas spills x[])
      trsm_strip<16>(x, tile == 0 ? seq : seq + 16, tile == 0 ? 8 : ns1, W8s + tile * 1024, lane, acquire, release);
#pragma unroll
      for (int ct = 0; ct < 16; ++ct) {
        ssq = fma(x[ct][0], x[ct][0], ssq);
        ssq = fma(x[ct][1], x[ct][1], ssq);
      }
      if (tile == 0 && nt == 2) {
#pragma unroll
        for (int ct = 0; ct < 16; ++ct)
          *reinterpret_cast<double2*>(Vs + row * PS_VLD + 8 * ct + 2 * tq) = make_double2(x[ct][0], x[ct][1]);
        __syncwarp();                     // V_0 rows are read back by this warp only
      }
    }
    fa += __shfl_xor_sync(FULLMASK, fa, 1); fa += __shfl_xor_sync(FULLMASK, fa, 2);
    ua += __shfl_xor_sync(FULLMASK, ua, 1); ua += __shfl_xor_sync(FULLMASK, ua, 2);
    ssq += __shfl_xor_sync(FULLMASK, ssq, 1); ssq += __shfl_xor_sync(FULLMASK, ssq, 2);
    const long gp = (long)pt * TB + row;
    if (tq == 0 && gp < A.npts) epilogue_point(A.epi, gp, fa, ua, ssq, zero_ss, any);
  }
  if (A.epi.stats) {
    if (zero_ss) atomicAdd(reinterpret_cast<unsigned long long*>(A.epi.stats), (unsigned long long)zero_ss);
    if (any) atomicOr(reinterpret_cast<unsigned long long*>(A.epi.stats + 1), 1ull);
  }
}

constexpr size_t PS_SMEM = ((size_t)TB * PS_VLD + PS_NST * SLAB_ELEMS + 2048 + PS_MAXD * 256 + 512 + 2 * PS_MAXD +
                            TB * (PS_MAXD + 1)) * 8 + 2 * PS_NST * 8;
static_assert(PS_SMEM <= 227 * 1024, "predict_small_kernel shared memory");

// ====================================================================== launchers
static inline size_t corr_smem_bytes(int d) {
  const int lda = d | 1;
  return (size_t)(((TB * lda + 1) & ~1) + d * TB + 4 * d) * sizeof(double);
}

#define KB_LAUNCH_CHECK()                          \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return e__;            \
  } while (0)

// ================================================================== leave-one-out (LOOCV)
// krigloocv2.py:7-34 refits the model n times on n-1 samples at fixed theta and predicts the
// left-out sample.  With A = Psi + eps I = L L^T, alpha = A^-1 (y - F beta), w = A^-1 F that is
// the closed form (Dubrule 1983)   y_loo[i] = y[i] - alpha[i] / (A^-1_ii - w[i]^2 / (F^T A^-1 F)),
// and A^-1_ii = ||L^-1 e_i||^2 comes out of the predictor's own variance machinery when the
// "correlation vectors" are the unit vectors: identity_chunk_kernel builds them, the panel kernel
// (pred mode) solves them, loo_kernel finishes.
__global__ void identity_chunk_kernel(double* chunk, int nt, long pt0) {
  // chunk tile (pt, j): rows = points pt0*128 + pt*128 + r, columns = samples j*128 + c
  double* T = chunk + ((size_t)blockIdx.x * nt + blockIdx.y) * TILE_ELEMS;
  const bool diag = (pt0 + blockIdx.x) == blockIdx.y;
  for (int idx = threadIdx.x; idx < TILE_ELEMS; idx += blockDim.x) T[idx] = 0.0;
  __syncthreads();
  if (diag)
    for (int r = threadIdx.x; r < TB; r += blockDim.x) T[tile_off(r, r)] = 1.0;
}

// fast_sigma2 <= 0: krigloocv2.py (the closed form of n refits).  fast_sigma2 > 0: krigloocv.py ('fast'): the bordered
// matrix B = [[sigma^2 Psi, f], [f^T, 1]] (a ONE in the corner, Psi without nugget -- the installed factor must be that
// of Psi itself) and LOO_i = y_i - (B^-1[:n,:n] y)_i / B^-1_ii.  With P = Psi^-1, w = P f, alpha = P (y - f beta),
// s = 1 - f^T P f / sigma^2:  (B^-1 y)_i = [alpha_i + beta w_i + w_i (beta f^T P f / sigma^2) / s] / sigma^2,
// B^-1_ii = [P_ii + w_i^2 / (sigma^2 s)] / sigma^2.
__global__ void loo_kernel(const double* y, const double* alpha, const double* w, const double* ssq, double btb,
                           int n, long p0, int np, double* out, double fast_sigma2, double beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long gi = p0 + i;
  if (gi >= n) return;
  if (fast_sigma2 > 0.0) {
    const double s = 1.0 - btb / fast_sigma2;
    const double num = alpha[gi] + beta * w[gi] + w[gi] * ((beta * btb / fast_sigma2) / s);
    const double den = ssq[i] + w[gi] * (w[gi] / (fast_sigma2 * s));
    out[gi] = y[gi] - num / den;
    return;
  }
  const double q = ssq[i] - w[gi] * (w[gi] / btb);
  out[gi] = y[gi] - alpha[gi] / q;
}

// ================================================== AK-MCS reductions (akmcs.py:267-297)
// From G = pred and sigma_G = s of a Monte-Carlo population: U = |G| / s with its minimum and
// first arg-min (lfucalc), #(G <= 0) (pfcalc), #(G - 1.96 s <= 0) and #(G + 1.96 s <= 0)
// (stopcrit).  One partial record per block; the handful of records is combined on the host.
struct AkPartial {
  double min_u;
  long long arg;
  unsigned long long c0, cm, cp;
};

__device__ __forceinline__ bool ak_better(double u, long long i, double bu, long long bi) {
  // np.argmin order: NaN first, then smaller value, ties -> smaller index
  const bool un = u != u, bn = bu != bu;
  if (un != bn) return un;
  if (un) return i < bi;
  return u < bu || (u == bu && i < bi);
}

__global__ void __launch_bounds__(256)
akmcs_reduce_kernel(const double* __restrict__ G, const double* __restrict__ S, long n, long base, AkPartial* out) {
  double bu = INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  unsigned long long c0 = 0, cm = 0, cp = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double g = G[i], sg = S[i];
    const double u = fabs(g) / sg;
    if (ak_better(u, base + i, bu, bi)) { bu = u; bi = base + i; }
    c0 += g <= 0.0;
    cm += (g - 1.96 * sg) <= 0.0;
    cp += (g + 1.96 * sg) <= 0.0;
  }
  __shared__ double su[256];
  __shared__ long long si[256];
  __shared__ unsigned long long sc[3][256];
  const int t = threadIdx.x;
  su[t] = bu; si[t] = bi; sc[0][t] = c0; sc[1][t] = cm; sc[2][t] = cp;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) {
      if (ak_better(su[t + o], si[t + o], su[t], si[t])) { su[t] = su[t + o]; si[t] = si[t + o]; }
      sc[0][t] += sc[0][t + o]; sc[1][t] += sc[1][t + o]; sc[2][t] += sc[2][t + o];
    }
    __syncthreads();
  }
  if (t == 0) {
    AkPartial p;
    p.min_u = su[0]; p.arg = si[0]; p.c0 = sc[0][0]; p.cm = sc[1][0]; p.cp = sc[2][0];
    out[blockIdx.x] = p;
  }
}

cudaError_t launch_akmcs_reduce(const double* G, const double* S, long n, long base, void* out, int nblocks, cudaStream_t st) {
  akmcs_reduce_kernel<<<nblocks, 256, 0, st>>>(G, S, n, base, static_cast<AkPartial*>(out));
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

// ===================================== one-sample append to a factor
// Kriging believer (optim_tools/MOBO.py:344-356) and the sequential updates of AK-MCS / SOBO
// (reliability_analysis/akmcs.py:127-139, optim_tools/SOBO.py:137-145) call likelihood(Theta, KrigInfo, 'all')
// again after ONE sample was stacked under X and y: the Cholesky factor of the bordered matrix is the old
// factor with one more row  [v^T, l],  v = L^-1 psi_new,  l = sqrt(psi_nn + eps - v.v): one forward solve
// (trsv kernels above; the panel kernel in prediction mode took 5 ms for a single point at n = 4000).
//
// append_extract_kernel: row `n_prev` of the freshly assembled bordered matrix (tile row ti_new, local row
// r_new) as a plain zero-padded vector of length nt_prev * 128, and its diagonal entry; one CTA per tile.
__global__ void __launch_bounds__(TB)
append_extract_kernel(const double* __restrict__ tiles, int n_prev, int nt_prev, double* __restrict__ vec,
                      double* __restrict__ dnn) {
  const int tj = blockIdx.x, ti_new = n_prev / TB, r_new = n_prev % TB, c = threadIdx.x;
  // psi_nn + eps exactly as the assembly wrote it (the old factor is copied over this tile next)
  if (tj == 0 && c == 0) *dnn = tiles[(size_t)tri_index(ti_new, ti_new) * TILE_ELEMS + tile_off(r_new, r_new)];
  const double* S = tiles + (size_t)tri_index(ti_new, tj) * TILE_ELEMS;
  vec[tj * TB + c] = (tj * TB + c < n_prev) ? S[tile_off(r_new, c)] : 0.0;
}

// append_row_kernel: one CTA.  `tiles` already holds the old factor (copied over the assembled matrix except
// for its new diagonal entry, kept aside by append_extract_kernel).
struct AppendArgs {
  double* tiles;           // bordered matrix / new factor, tile storage
  const double* v;         // [nt_prev * 128]  L^-1 psi_new
  int n_prev, nt_prev;
  int nq, npad_prev, npad;
  const double* Zprev;     // [nq][npad_prev]  L^-1 [y F] of the old fit
  const double* R;         // [nq][npad]       [y F]^T of the bordered model
  double* Z;               // [nq][npad]
  const double* logdet_prev;
  double* logdet;
  int* info;
  const double* dnn;       // diagonal entry of the bordered matrix
};
__global__ void __launch_bounds__(256)
append_row_kernel(AppendArgs A) {
  __shared__ double red[8];
  __shared__ double s_l;
  const int n = A.n_prev, ti = n / TB, r = n % TB;
  double* D = A.tiles + (size_t)tri_index(ti, ti) * TILE_ELEMS;
  const double dnn = *A.dnn;
  double ss = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double v = A.v[c];
    ss = fma(v, v, ss);
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    const double piv = dnn - ss;
    double l = 1.0;
    if (piv > 0.0) l = sqrt(piv);
    else *A.info = n + 1;                        // first non-positive pivot, 1-based (as the factorisation reports)
    s_l = l;
    D[tile_off(r, r)] = l;
    *A.logdet = *A.logdet_prev + log(l);
  }
  __syncthreads();
  const double l = s_l;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double v = A.v[c];
    A.tiles[(size_t)tri_index(ti, c / TB) * TILE_ELEMS + tile_off(r, c % TB)] = v;
  }
  // z_new = ([y F]_new - v . Z_prev) / l   (forward substitution of likelihood_func.py:183-193, last row)
  for (int q = 0; q < A.nq; ++q) {
    const double* zp = A.Zprev + (size_t)q * A.npad_prev;
    double* zn = A.Z + (size_t)q * A.npad;
    double dot = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
      const double v = A.v[c];
      const double z = zp[c];
      dot = fma(v, z, dot);
      zn[c] = z;
    }
    dot = block_sum(dot, red);
    for (int c = n + 1 + threadIdx.x; c < A.npad; c += blockDim.x) zn[c] = 0.0;
    if (threadIdx.x == 0) zn[n] = (A.R[(size_t)q * A.npad + n] - dot) / l;
  }
}

cudaError_t launch_append_extract(const double* tiles, int n_prev, int nt_prev, double* vec, double* dnn,
                                  cudaStream_t st) {
  append_extract_kernel<<<nt_prev, TB, 0, st>>>(tiles, n_prev, nt_prev, vec, dnn);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t launch_append_row(double* tiles, const double* v, int n_prev, int nt_prev, int nq, int npad_prev,
                              int npad, const double* Zprev, const double* R, double* Z, const double* logdet_prev,
                              double* logdet, int* info, const double* dnn, cudaStream_t st) {
  AppendArgs a;
  a.dnn = dnn;
  a.tiles = tiles; a.v = v; a.n_prev = n_prev; a.nt_prev = nt_prev; a.nq = nq; a.npad_prev = npad_prev;
  a.npad = npad; a.Zprev = Zprev; a.R = R; a.Z = Z; a.logdet_prev = logdet_prev; a.logdet = logdet; a.info = info;
  append_row_kernel<<<1, 256, 0, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_identity_chunk(double* chunk, int tiles, int nt, long pt0, cudaStream_t st) {
  identity_chunk_kernel<<<dim3(tiles, nt), 256, 0, st>>>(chunk, nt, pt0);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_loo(const double* y, const double* alpha, const double* w, const double* ssq, double btb, int n,
                       long p0, int np, double* out, cudaStream_t st, double fast_sigma2, double beta) {
  loo_kernel<<<(np + 255) / 256, 256, 0, st>>>(y, alpha, w, ssq, btb, n, p0, np, out, fast_sigma2, beta);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}



// dynamic-smem opt-in is per device (context): remember what was set for each device
static size_t g_corr_attr[64] = {0};
static bool g_chol_attr[64] = {false};
static inline int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
constexpr size_t DIAG_SMEM = ((size_t)NSTAGE * STAGE_ELEMS + DIAG_EXTRA_ELEMS) * 8 + NBARRIERS * 8 + 16;
constexpr size_t PANEL_SMEM = ((size_t)PANEL_EXTRA_OFF + 8 * TLD + TB + 16 + 64) * 8 + NBARRIERS * 8 + 16;
static_assert(DIAG_SMEM <= 227 * 1024 && PANEL_SMEM <= 227 * 1024, "shared memory budget");
static_assert((size_t)TB * LDP * 8 + 16 <= (size_t)NSTAGE * STAGE_ELEMS * 8, "POTF2 scratch must fit in the stage buffers");

cudaError_t launch_corr(const CorrArgs& a, int grid_x, int grid_y, cudaStream_t st) {
  const size_t sm = corr_smem_bytes(a.spec.d);
  if (sm > 227 * 1024) return cudaErrorInvalidValue;
  const int dev = cur_dev();
  if (sm > g_corr_attr[dev]) {
    const int want = (int)(sm > 48 * 1024 ? sm : 48 * 1024);
    const void* kernels[] = {(const void*)corr_tile_kernel<0, CORR_ASSEMBLE>, (const void*)corr_tile_kernel<1, CORR_ASSEMBLE>,
                             (const void*)corr_tile_kernel<4, CORR_ASSEMBLE>, (const void*)corr_tile_kernel<0, CORR_PREDICT>,
                             (const void*)corr_tile_kernel<1, CORR_PREDICT>,  (const void*)corr_tile_kernel<2, CORR_PREDICT>,
                             (const void*)corr_tile_kernel<4, CORR_PREDICT>};
    for (const void* k : kernels) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
      if (e != cudaSuccess) return e;
    }
    g_corr_attr[dev] = sm;
  }
  // Gaussian fast path: the likelihood's assembly for plain distances only (numpy's operations bit for bit); prediction
  // whenever the caller provides the per-dimension scales -- a KPLS-Gaussian correlation is a Gaussian in d dimensions with
  // weights sum_c pls_kc^2 / theta_c^2 (kernelfunc.py:42-47), 2 d FP64 instructions per pair instead of d (2 + 3 h)
  const bool fast = a.spec.nkernel == 1 && a.spec.kid[0] == K_GAUSS &&
                    (a.mode == CORR_PREDICT ? (a.pscale != nullptr && !KB200_PREDICT_EXACT_DIV) || !a.spec.kpls : !a.spec.kpls);
  // One kernel per class of d where that measured faster (r1s A/B, profiles/r1s_ab_assembly_pairing.log): d < 8 both modes
  // (prediction at d = 6 2.09 -> 1.91 ms per 2e6 points), 8 <= d < 16 prediction (8.11 -> 7.91 ms); the kernel holding all
  // three summation schemes (4) elsewhere: specialised, ptxas spills in the assembly at 8 <= d < 16 (1.53 -> 1.65 ms) and
  // for d >= 16 (C4 11.8 -> 14.2 ms).
  const int d = a.spec.d;
  const dim3 grid(grid_x, grid_y);
#define KB_CORR_LAUNCH(DC, MODE) corr_tile_kernel<DC, MODE><<<grid, CORR_THREADS, sm, st>>>(a)
  if (a.mode == CORR_ASSEMBLE) {
    if (!fast) KB_CORR_LAUNCH(0, CORR_ASSEMBLE);
    else if (d < 8) KB_CORR_LAUNCH(1, CORR_ASSEMBLE);
    else KB_CORR_LAUNCH(4, CORR_ASSEMBLE);
  } else {
    if (!fast) KB_CORR_LAUNCH(0, CORR_PREDICT);
    else if (d < 8) KB_CORR_LAUNCH(1, CORR_PREDICT);
    else if (d < 16) KB_CORR_LAUNCH(2, CORR_PREDICT);
    else KB_CORR_LAUNCH(4, CORR_PREDICT);
  }
#undef KB_CORR_LAUNCH
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

static cudaError_t chol_attrs() {
  const int dev = cur_dev();
  if (g_chol_attr[dev]) return cudaSuccess;
  cudaError_t e;
  e = cudaFuncSetAttribute(chol_diag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chol_diag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chol_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chol_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM);
  if (e != cudaSuccess) return e;
  g_chol_attr[dev] = true;
  return cudaSuccess;
}

cudaError_t launch_chol_diag(const CholArgs& a, int j, int nmat, cudaStream_t st) {
  cudaError_t e = chol_attrs();
  if (e != cudaSuccess) return e;
  if (a.naive) chol_diag_kernel<true><<<nmat, CH_THREADS, DIAG_SMEM, st>>>(a, j);
  else chol_diag_kernel<false><<<nmat, CH_THREADS, DIAG_SMEM, st>>>(a, j);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_chol_panel(const CholArgs& a, int j, int grid_x, int grid_y, cudaStream_t st) {
  cudaError_t e = chol_attrs();
  if (e != cudaSuccess) return e;
  if (grid_x <= 0 || grid_y <= 0) return cudaSuccess;
  if (a.naive) chol_panel_kernel<true><<<dim3(grid_x, grid_y), CH_THREADS, PANEL_SMEM, st>>>(a, j);
  else chol_panel_kernel<false><<<dim3(grid_x, grid_y), CH_THREADS, PANEL_SMEM, st>>>(a, j);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_w8_from_tiles(const CholArgs& a, int nmat, cudaStream_t st) {
  w8_from_tiles_kernel<<<dim3(a.nt, nmat), 128, 0, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_finalize(const FinArgs& a, int nmat, cudaStream_t st) {
  lik_finalize_kernel<<<nmat, 256, 0, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

static bool g_solve_attr[64] = {false};
static cudaError_t solve_attrs() {
  const int dev = cur_dev();
  if (g_solve_attr[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(trsv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM);
  if (e != cudaSuccess) return e;
  g_solve_attr[dev] = true;
  return cudaSuccess;
}

static cudaError_t launch_trsv(const BackArgs& a, bool transposed, cudaStream_t st) {
  cudaError_t e = solve_attrs();
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(a.work, a.rhs, (size_t)a.nq * a.npad * 8, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  for (int s = 0; s < a.nt; ++s) {
    const int j = transposed ? a.nt - 1 - s : s;
    trsv_diag_kernel<<<1, CH_THREADS, SOLVE_SMEM, st>>>(a, j, transposed ? 1 : 0);
    KB_LAUNCH_CHECK();
    const int rest = transposed ? j : a.nt - 1 - j;
    if (rest > 0) {
      trsv_update_kernel<<<rest, CH_THREADS, 0, st>>>(a, j, transposed ? 1 : 0);
      KB_LAUNCH_CHECK();
    }
  }
  return cudaSuccess;
}

cudaError_t launch_backsolve(const BackArgs& a, cudaStream_t st) { return launch_trsv(a, true, st); }
cudaError_t launch_forwardsolve(const BackArgs& a, cudaStream_t st) { return launch_trsv(a, false, st); }

cudaError_t launch_decode(const DecodeArgs& a, cudaStream_t st) {
  decode_kernel<<<(int)((a.P + 127) / 128), 128, 0, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_import_U(const double* U, int n, int nt, double* tiles, cudaStream_t st) {
  import_U_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(U, n, nt, tiles);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_export(const double* tiles, int n, double* out, int what, cudaStream_t st) {
  export_kernel<<<148 * 4, 256, 0, st>>>(tiles, n, out, what);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_add_diag(double* tiles, int n, const double* eps_ptr, cudaStream_t st) {
  add_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(tiles, n, eps_ptr);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_epilogue(const EpiArgs& a, cudaStream_t st) {
  long blocks = (a.npts + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  predict_epilogue_kernel<<<(int)blocks, 256, 0, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

// ===================================================================== FP64 probes
// Register-resident DMMA / DFMA loops: the empirical FP64 roof of this chip (there is no
// FP64 figure in MEASURED_PEAKS.json).  16 independent accumulator tiles per warp, like the
// GEMM core, so the number is the ceiling of that instruction mix.
__global__ void __launch_bounds__(CH_THREADS, 1) fp64_probe_kernel(int iters, int mode, double* sink) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        acc[i][0] = fma(a, b, acc[i][0]);
        acc[i][1] = fma(b, a, acc[i][1]);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

cudaError_t launch_predict_small(const PredSmallArgs& a, int sm_count, cudaStream_t st) {
  static bool attr[64] = {false};
  const int dev = cur_dev();
  if (!attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(predict_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const long tiles = (a.npts + TB - 1) / TB;
  const int grid = (int)(tiles < sm_count ? tiles : sm_count);
  predict_small_kernel<<<grid, CH_THREADS, PS_SMEM, st>>>(a);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

bool predict_small_supported(int n, int d, int nkernel, int kid0, int kpls) {
  return n <= 2 * TB && d <= PS_MAXD && nkernel == 1 && kid0 == K_GAUSS && (!kpls || !KB200_PREDICT_EXACT_DIV);
}

// scales of the pre-scaled prediction coordinates (kb200_corr.cuh:corr_quad_gauss_scaled)
__global__ void pred_scale_kernel(const double* __restrict__ cand, const double* __restrict__ pls2, int d, int ntheta, int kpls,
                                  double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d) return;
  if (!kpls) {
    out[k] = cand[ntheta + k];                       // RN(1 / theta_k)
  } else {
    double s = 0.0;                                  // sum_c pls_kc^2 * RN(1 / theta_c^2)
    for (int c = 0; c < ntheta; ++c) s = fma(pls2[(size_t)k * ntheta + c], cand[3 * ntheta + c], s);
    out[k] = sqrt(s);
  }
}
cudaError_t launch_pred_scale(const double* cand_row, const double* pls2, int d, int ntheta, int kpls, double* out, cudaStream_t st) {
  pred_scale_kernel<<<(d + 127) / 128, 128, 0, st>>>(cand_row, pls2, d, ntheta, kpls, out);
  return cudaGetLastError();
}

// debug / parity: the prediction kernels' exp (kb200_corr.cuh:exp_nonpos; mode 0) or libm's (mode 1), element-wise
__global__ void debug_exp_kernel(const double* __restrict__ x, long n, int mode, double* __restrict__ y) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = mode == 0 ? exp_nonpos(x[i]) : exp(x[i]);
}
cudaError_t launch_debug_exp(const double* x, long n, int mode, double* y, cudaStream_t st) {
  debug_exp_kernel<<<296, 256, 0, st>>>(x, n, mode, y);
  return cudaGetLastError();
}

cudaError_t launch_fp64_probe(int blocks, int iters, int mode, double* sink, cudaStream_t st) {
  fp64_probe_kernel<<<blocks, CH_THREADS, 0, st>>>(iters, mode, sink);
  KB_LAUNCH_CHECK();
  return cudaSuccess;
}

}  // namespace kb200
