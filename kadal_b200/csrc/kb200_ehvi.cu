// kadal_b200 — 2-objective expected hyper-volume improvement on the device.
//
// Replaces kadal/optim_tools/ehvi/exi2d.py:7-84 (with exipsi.py:4-6, gaussfcn.py:4-10,
// hvolume2d.py:3-14) as called per candidate by EHVIcomputation.py:57 and :169-175.
// The reference walks the (k+1)(k+2)/2 staircase cells of the sorted Pareto front for every
// candidate in Python (and rebuilds the dominated set of every cell with an O(k) loop: O(k^3) per
// candidate).  The cell geometry depends only on the front:
//   ehvi_sort_kernel    sorts the front (objective 1 ascending; objective-2 list ascending)
//   ehvi_table_kernel   sPlus of every cell (hvolume2d of the points the cell's upper corner
//                       dominates), one thread per cell
//   ehvi_score_kernel   one CTA per candidate: the k+2 Gaussian cdf/pdf values per objective go to
//                       shared memory once, then every cell is a handful of multiplies; thread b
//                       accumulates column b over the rows in order and the column sums are
//                       combined in numpy's pairwise order, i.e. np.sum(np.sum(c, axis=0)).
// Every product / sum is rounded separately (no FMA contraction) in the reference's order, so a
// cell differs from the reference only through the last ulp of erf / exp.
#include <cmath>
#include "kb200_common.cuh"
#include "kb200_kernels.h"

namespace kb200 {

// ranks with ties broken by position (a stable sort)
__global__ void ehvi_sort_kernel(const double* __restrict__ P, int k, double* __restrict__ S, double* __restrict__ c2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    const double a = P[2 * i], b = P[2 * i + 1];
    int ra = 0, rb = 0;
    for (int j = 0; j < k; ++j) {
      const double aj = P[2 * j], bj = P[2 * j + 1];
      ra += (aj < a) || (aj == a && j < i);
      rb += (bj < b) || (bj == b && j < i);
    }
    S[2 * ra] = a;
    S[2 * ra + 1] = b;
    c2[rb] = b;
  }
}

// cell (a, b) = reference loop indices (i, j) = (a - 1, b - 1); valid when b <= k - a.
// c1[t] = S[t][0].  fmax1 = a ? c1[k-a] : r0, fmax2 = b ? c2[k-b] : r1,
// cU1 = b == k ? r0 : c1[b], cU2 = a == k ? r1 : c2[a]            (exi2d.py:22-46)
__global__ void ehvi_table_kernel(const double* __restrict__ S, const double* __restrict__ c2, int k, double r0, double r1,
                                  double* __restrict__ splus) {
  const int k1 = k + 1;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < k1 * k1; idx += gridDim.x * blockDim.x) {
    const int a = idx / k1, b = idx - a * k1;
    double h = 0.0;
    if (b <= k - a) {
      const double f1 = a ? S[2 * (k - a)] : r0, f2 = b ? c2[k - b] : r1;
      const double cU1 = b == k ? r0 : S[2 * b], cU2 = a == k ? r1 : c2[a];
      bool first = true;
      double prev = 0.0;
      for (int m = 0; m < k; ++m) {        // hvolume2d over the dominated points, objective 1 ascending
        const double s0 = S[2 * m], s1 = S[2 * m + 1];
        if (cU1 <= s0 && cU2 <= s1) {
          const double w = __dsub_rn(f1, s0);
          h = first ? __dmul_rn(w, __dsub_rn(f2, s1)) : __dadd_rn(h, __dmul_rn(w, __dsub_rn(prev, s1)));
          prev = s1;
          first = false;
        }
      }
    }
    splus[idx] = h;
  }
}

// numpy's pairwise summation (loops_utils.h.src) of a contiguous vector
__device__ double np_pairwise(const double* a, int n) {
  if (n <= 128) return np_sum_block(0, n, [&](int i) { return a[i]; });
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}

constexpr int EHVI_THREADS = 128;

__global__ void __launch_bounds__(EHVI_THREADS)
ehvi_score_kernel(EhviArgs A) {
  extern __shared__ double sm[];
  const int k = A.k, k1 = k + 1, k2 = k + 2;
  double* cdf1 = sm;            // values at B1 = [-inf, c1[0..k-1], r0]
  double* pdf1 = cdf1 + k2;
  double* cdf2 = pdf1 + k2;     // values at B2 = [-inf, c2[0..k-1], r1]
  double* pdf2 = cdf2 + k2;
  double* col = pdf2 + k2;      // [k+1] column sums
  const double INV_SQRT_2PI = 0x1.9884533d43651p-2;   // 1/np.sqrt(2*np.pi)
  const double SQRT2 = 0x1.6a09e667f3bcdp+0;          // np.sqrt(2)
  const long i = blockIdx.x;
  const double m0 = A.mu0[i * A.inc], m1 = A.mu1[i * A.inc], s0 = A.s0[i * A.inc], s1 = A.s1[i * A.inc];
  for (int t = threadIdx.x; t < 2 * k2; t += EHVI_THREADS) {
    const int dim = t >= k2, u = dim ? t - k2 : t;
    const double bnd = u == 0 ? -INFINITY : (u == k1 ? (dim ? A.r1 : A.r0) : (dim ? A.c2[u - 1] : A.S[2 * (u - 1)]));
    const double z = __ddiv_rn(__dsub_rn(bnd, dim ? m1 : m0), dim ? s1 : s0);
    const double cdf = __dmul_rn(0.5, __dadd_rn(1.0, erf(__ddiv_rn(z, SQRT2))));          // gaussfcn.py:8-10
    const double pdf = __dmul_rn(INV_SQRT_2PI, exp(-__dmul_rn(z, z) / 2.0));               // gaussfcn.py:4-6
    (dim ? cdf2 : cdf1)[u] = cdf;
    (dim ? pdf2 : pdf1)[u] = pdf;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < k1; b += EHVI_THREADS) {
    const double f2 = b ? A.c2[k - b] : A.r1;
    const double G1 = __dsub_rn(cdf1[b + 1], cdf1[b]);
    const double pU1 = __dmul_rn(s0, pdf1[b + 1]), pL1 = __dmul_rn(s0, pdf1[b]);
    const double cU1 = cdf1[b + 1], cL1 = cdf1[b];
    const double d2 = __dsub_rn(f2, m1);
    double acc = 0.0;
    for (int a = 0; a <= k - b; ++a) {
      const double f1 = a ? A.S[2 * (k - a)] : A.r0;
      const double d1 = __dsub_rn(f1, m0);
      // exipsi(fmax, cU) - exipsi(fmax, cL)                               exipsi.py:5, exi2d.py:65-67
      const double Psi1 = __dsub_rn(__dadd_rn(pU1, __dmul_rn(d1, cU1)), __dadd_rn(pL1, __dmul_rn(d1, cL1)));
      const double Psi2 = __dsub_rn(__dadd_rn(__dmul_rn(s1, pdf2[a + 1]), __dmul_rn(d2, cdf2[a + 1])),
                                    __dadd_rn(__dmul_rn(s1, pdf2[a]), __dmul_rn(d2, cdf2[a])));
      const double G2 = __dsub_rn(cdf2[a + 1], cdf2[a]);
      const double v = __dsub_rn(__dmul_rn(Psi1, Psi2), __dmul_rn(__dmul_rn(A.splus[(size_t)a * k1 + b], G1), G2));
      acc = __dadd_rn(acc, v < 0.0 ? 0.0 : v);                             // exi2d.py:75-81 (NaN stays NaN)
    }
    col[b] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) A.out[i] = A.sign * np_pairwise(col, k1);
}

cudaError_t launch_ehvi_tables(const double* P, int k, double r0, double r1, double* S, double* c2, double* splus,
                               cudaStream_t st) {
  if (k > 0) {
    ehvi_sort_kernel<<<(k + 127) / 128, 128, 0, st>>>(P, k, S, c2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const long cells = (long)(k + 1) * (k + 1);
  ehvi_table_kernel<<<(int)((cells + 127) / 128), 128, 0, st>>>(S, c2, k, r0, r1, splus);
  return cudaGetLastError();
}

cudaError_t launch_ehvi_score(const EhviArgs& a, long npop, cudaStream_t st) {
  if (npop <= 0) return cudaSuccess;
  const size_t sm = (size_t)(5 * (a.k + 2)) * sizeof(double);
  if (sm > 48 * 1024) return cudaErrorInvalidValue;
  ehvi_score_kernel<<<(unsigned)npop, EHVI_THREADS, sm, st>>>(a);
  return cudaGetLastError();
}

}  // namespace kb200
