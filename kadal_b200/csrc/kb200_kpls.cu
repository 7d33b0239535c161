// kadal_b200 — KPLS correlation assembly from cached projected-distance planes.
//
// kernelfunc.py:41-47 (gaussian) / :58-64 (exponential) with a PLS rotation compute
//     t_c(i,j) = sum_k g(x_ik - x_jk) * coef[k][c],   g = (.)^2 with pls^2  |  |.| with pls
//     Psi_ij   = exp(-0.5 * sum_c t_c / theta_c^2)    |  exp(-sum_c t_c / theta_c)
// The h planes t_c (h = n_princomp, 2..4 in practice) do NOT depend on the hyper-parameters:
// during Kriging.train the reference recomputes the (n, n, d) distance stack and its dgemm with
// pls for every one of the thousands of likelihood calls (the 12.8 GB temporary of C5).  Here the
// planes are built once per model (kpls_planes_kernel: d-long FMA chains, the exact operation
// order of the on-the-fly path in kb200_corr.cuh, so results are bit-identical to it) and kept
// in HBM in the factor's own tile layout; a candidate's Psi is then h divisions, a sum and one exp
// per entry (kpls_assemble_kernel): an HBM-streaming kernel, 8 (h + 1) bytes per entry, with the
// candidate index as the fast grid dimension so that the CTAs that share a plane tile run
// together and hit it in L2.
#include "kb200_kernels.h"

namespace kb200 {

constexpr int KP_THREADS = 256;

// one CTA = one lower tile (ti, tj); all h planes of one distance flavour
__global__ void __launch_bounds__(KP_THREADS, 1)
kpls_planes_kernel(const double* __restrict__ X, int n, int d, int h, const double* __restrict__ coef, int absmode,
                   double* __restrict__ planes, long tiles_per_mat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lda = d | 1;
  double* As = reinterpret_cast<double*>(smem_raw);          // [128][lda]
  double* Bt = As + ((TB * lda + 1) & ~1);                   // [d][128]
  double* cf = Bt + d * TB;                                  // [d][h]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, ch = lane & 3;
  const int t = blockIdx.x;
  int ti = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  for (int idx = tid; idx < TB * d; idx += KP_THREADS) {
    const int r = idx / d, k = idx - r * d;
    const int gr = ti * TB + r, gc = tj * TB + r;
    As[r * lda + k] = gr < n ? X[(size_t)gr * d + k] : 0.0;
    Bt[k * TB + r] = gc < n ? X[(size_t)gc * d + k] : 0.0;
  }
  for (int idx = tid; idx < d * h; idx += KP_THREADS) cf[idx] = coef[idx];
  __syncthreads();
  for (int c0 = 0; c0 < h; c0 += 4) {
    const int nc = h - c0 < 4 ? h - c0 : 4;
#pragma unroll 1
    for (int rgi = 0; rgi < 2; ++rgi) {
      const int row = 8 * (warp + 8 * rgi) + g;
      const double* a = As + row * lda;
#pragma unroll 1
      for (int s = 0; s < NSLAB; ++s) {
        const int col0 = s * KS + 4 * ch;
        double acc[4][4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[cc][e] = 0.0;
        for (int k = 0; k < d; ++k) {
          const double ak = a[k];
          const Quad b = load_b4(Bt + col0, k, TB);
          double gv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const double df = __dsub_rn(ak, b.v[e]);
            gv[e] = absmode ? fabs(df) : __dmul_rn(df, df);
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc < nc) {
              const double c = cf[k * h + c0 + cc];
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[cc][e] = fma(gv[e], c, acc[cc][e]);   // dgemm inner product, order k
            }
          }
        }
        const int off = s * SLAB_ELEMS + row * KS + ((ch ^ (row & 3)) << 2);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc < nc) {
            double2* dst = reinterpret_cast<double2*>(planes + ((size_t)(c0 + cc) * tiles_per_mat + t) * TILE_ELEMS + off);
            dst[0] = make_double2(acc[cc][0], acc[cc][1]);
            dst[1] = make_double2(acc[cc][2], acc[cc][3]);
          }
        }
      }
    }
  }
}

// grid (P, tiles_per_mat): candidate m, lower tile t.  Thread -> 16 quads of the tile, linear in
// memory (a quad is one 32-byte chunk of a slab line), so every load and store is coalesced.
__global__ void __launch_bounds__(KP_THREADS)
kpls_assemble_kernel(KplsAsmArgs A) {
  __shared__ double prm[4 * 8 + MAXK + 2];
  const int h = A.h, nk = A.nkernel;
  const long m = blockIdx.x;
  const int t = blockIdx.y;
  const double* cand = A.cand + m * A.cand_stride;
  if (threadIdx.x < A.cand_stride) prm[threadIdx.x] = cand[threadIdx.x];
  int ti = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  __syncthreads();
  const double* theta = prm;
  const double* rtheta = prm + h;
  const double* theta2 = prm + 2 * h;
  const double* rtheta2 = prm + 3 * h;
  const double* wg = prm + 4 * h;
  const double eps = prm[4 * h + nk];
  double* out = A.out + ((size_t)m * A.tiles_per_mat + t) * TILE_ELEMS;
  const size_t plane_stride = (size_t)A.tiles_per_mat * TILE_ELEMS;
  const double* sq = A.planes_sq ? A.planes_sq + (size_t)t * TILE_ELEMS : nullptr;
  const double* ab = A.planes_abs ? A.planes_abs + (size_t)t * TILE_ELEMS : nullptr;
#pragma unroll 1
  for (int q = threadIdx.x; q < TILE_ELEMS / 4; q += KP_THREADS) {
    const int chs = q & 3, row = (q >> 2) & (TB - 1), s = q >> 9;
    const int col0 = s * KS + ((chs ^ (row & 3)) << 2);
    const int grow = ti * TB + row, gc0 = tj * TB + col0;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const bool need = !(grow >= A.n || gc0 >= A.n || (ti == tj && col0 > row));
    if (need) {
      double Ssq[4] = {0.0, 0.0, 0.0, 0.0}, Sab[4] = {0.0, 0.0, 0.0, 0.0};
      for (int c = 0; c < h; ++c) {
        if (sq) {
          const double2 lo = *reinterpret_cast<const double2*>(sq + c * plane_stride + 4 * q);
          const double2 hi = *reinterpret_cast<const double2*>(sq + c * plane_stride + 4 * q + 2);
          const double t2 = theta2[c], r2 = rtheta2[c];
          Ssq[0] = __dadd_rn(Ssq[0], div_by_rcp(lo.x, t2, r2));
          Ssq[1] = __dadd_rn(Ssq[1], div_by_rcp(lo.y, t2, r2));
          Ssq[2] = __dadd_rn(Ssq[2], div_by_rcp(hi.x, t2, r2));
          Ssq[3] = __dadd_rn(Ssq[3], div_by_rcp(hi.y, t2, r2));
        }
        if (ab) {
          const double2 lo = *reinterpret_cast<const double2*>(ab + c * plane_stride + 4 * q);
          const double2 hi = *reinterpret_cast<const double2*>(ab + c * plane_stride + 4 * q + 2);
          const double t1 = theta[c], r1 = rtheta[c];
          Sab[0] = __dadd_rn(Sab[0], div_by_rcp(lo.x, t1, r1));
          Sab[1] = __dadd_rn(Sab[1], div_by_rcp(lo.y, t1, r1));
          Sab[2] = __dadd_rn(Sab[2], div_by_rcp(hi.x, t1, r1));
          Sab[3] = __dadd_rn(Sab[3], div_by_rcp(hi.y, t1, r1));
        }
      }
      for (int kk = 0; kk < nk; ++kk) {
        const bool gauss = A.kid[kk] == K_GAUSS;
        const double w = wg[kk];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double val = gauss ? exp(-0.5 * Ssq[e]) : exp(-Sab[e]);
          v[e] = __dadd_rn(v[e], __dmul_rn(w, val));
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gc = gc0 + e;
      if (grow >= A.n || gc >= A.n) v[e] = (grow == gc) ? 1.0 : 0.0;           // identity padding
      else if (grow == gc && A.add_eps) v[e] = __dadd_rn(v[e], eps);           // Psi + eye*eps
      else if (ti == tj && gc > grow) v[e] = 0.0;
    }
    double2* dst = reinterpret_cast<double2*>(out + 4 * q);
    dst[0] = make_double2(v[0], v[1]);
    dst[1] = make_double2(v[2], v[3]);
  }
}

cudaError_t launch_kpls_planes(const double* X, int n, int d, int h, const double* coef, int absmode, double* planes,
                               long tiles_per_mat, cudaStream_t st) {
  const int lda = d | 1;
  const size_t sm = (size_t)(((TB * lda + 1) & ~1) + d * TB + d * h) * sizeof(double);
  if (sm > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kpls_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sm > 48 * 1024 ? sm : 48 * 1024));
  if (e != cudaSuccess) return e;
  kpls_planes_kernel<<<(int)tiles_per_mat, KP_THREADS, sm, st>>>(X, n, d, h, coef, absmode, planes, tiles_per_mat);
  return cudaGetLastError();
}

cudaError_t launch_kpls_assemble(const KplsAsmArgs& a, int P, cudaStream_t st) {
  if (a.h > 7 || a.tiles_per_mat > 65535) return cudaErrorInvalidValue;
  kpls_assemble_kernel<<<dim3(P, (int)a.tiles_per_mat), KP_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace kb200
