// kadal_b200 — device-side Saltelli sums for Sobol indices.
//
// kadal/sensitivity_analysis/sobol_ind.py:45-141 predicts the surrogate on A, B and on every
// C_i = B with column i taken from A (:108-109; two columns for second order, :152-154) in 10 000
// row slices, keeps the (N,1) prediction vectors on the host and reduces them with np.sum.  Here the
// N x d matrices C_i are never materialised beyond one chunk (sobol_compose_kernel), the predictions
// stay on the device and only the sums  sum(ya), sum(ya^2), sum(ya*yc), sum(yb*yc)  come back.
// Reductions are deterministic: a fixed grid writes per-block partials (thread-sequential, then a
// shared-memory tree), one thread folds them into the running totals in block order.
#include "kb200_kernels.h"

namespace kb200 {

__global__ void sobol_compose_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                     const unsigned char* __restrict__ from_a, long n, int d, double* __restrict__ C) {
  const long total = n * d;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d);
    C[idx] = from_a[k] ? A[idx] : B[idx];
  }
}

// mode 0: (sum u*w, sum v*w)      mode 1: (sum u, sum u*u)
__global__ void __launch_bounds__(256)
sobol_dot_kernel(const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w, long n,
                 int mode, double* __restrict__ partial) {
  double s0 = 0.0, s1 = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      const double wi = w[i];
      s0 += u[i] * wi;
      s1 += v[i] * wi;
    } else {
      const double ui = u[i];
      s0 += ui;
      s1 += ui * ui;
    }
  }
  __shared__ double r0[256], r1[256];
  const int t = threadIdx.x;
  r0[t] = s0;
  r1[t] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) { r0[t] += r0[t + o]; r1[t] += r1[t + o]; }
    __syncthreads();
  }
  if (t == 0) { partial[blockIdx.x] = r0[0]; partial[gridDim.x + blockIdx.x] = r1[0]; }
}

__global__ void sobol_accum_kernel(const double* __restrict__ partial, int nblocks, double* acc0, double* acc1) {
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < nblocks; ++b) { s0 += partial[b]; s1 += partial[nblocks + b]; }
  *acc0 += s0;
  *acc1 += s1;
}

// Unscrambled Sobol points in Gray-code order, as scipy.stats.qmc.Sobol(scramble=False) and the reference's
// sobolsampling package generate them one after the other (x_i = x_{i-1} xor v[lowest zero bit of i-1]): in closed form
// x_i = xor of the direction numbers v[b] over the set bits b of gray(i) = i ^ (i >> 1), which needs no sequential state,
// so every thread makes its own point.  value = x * 2^-bits (exact), then optionally realval's samp * (ub - lb) + lb
// (samplingplan.py:37-51) without FMA contraction.  Row r of the call is sequence index i0 + r.  Column k < split goes to
// outA[r][k], column k >= split to outB[r][k - split] (the Saltelli A | B halves, sobol_ind.py:28-29); outB null: all
// `dims` columns into outA.
__global__ void __launch_bounds__(256)
sobol_points_kernel(const unsigned long long* __restrict__ dirnum, int bits, int dims, long long i0, long long n,
                    const double* __restrict__ lo, const double* __restrict__ hi, double* __restrict__ outA,
                    double* __restrict__ outB, int split) {
  const double scale = ldexp(1.0, -bits);
  const long long total = n * dims;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / dims;
    const int k = (int)(idx - r * dims);
    const unsigned long long i = (unsigned long long)(i0 + r);
    unsigned long long gray = i ^ (i >> 1), x = 0;
    const unsigned long long* v = dirnum + (size_t)k * bits;
    while (gray) {
      const int b = __ffsll((long long)gray) - 1;
      x ^= v[b];
      gray &= gray - 1;
    }
    double u = __dmul_rn((double)x, scale);
    if (lo) u = __dadd_rn(__dmul_rn(u, __dsub_rn(hi[k], lo[k])), lo[k]);
    if (outB == nullptr) outA[idx] = u;
    else if (k < split) outA[r * split + k] = u;
    else outB[r * (dims - split) + (k - split)] = u;
  }
}

cudaError_t launch_sobol_points(const unsigned long long* dirnum, int bits, int dims, long long i0, long long n,
                                const double* lo, const double* hi, double* outA, double* outB, int split, cudaStream_t st) {
  long long blocks = (n * dims + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sobol_points_kernel<<<(int)blocks, 256, 0, st>>>(dirnum, bits, dims, i0, n, lo, hi, outA, outB, split);
  return cudaGetLastError();
}

cudaError_t launch_sobol_compose(const double* A, const double* B, const unsigned char* from_a, long n, int d, double* C,
                                 cudaStream_t st) {
  long blocks = (n * d + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sobol_compose_kernel<<<(int)blocks, 256, 0, st>>>(A, B, from_a, n, d, C);
  return cudaGetLastError();
}

cudaError_t launch_sobol_dot(const double* u, const double* v, const double* w, long n, int mode, double* partial,
                             double* acc0, double* acc1, cudaStream_t st) {
  sobol_dot_kernel<<<SOBOL_BLOCKS, 256, 0, st>>>(u, v, w, n, mode, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sobol_accum_kernel<<<1, 1, 0, st>>>(partial, SOBOL_BLOCKS, acc0, acc1);
  return cudaGetLastError();
}

}  // namespace kb200
