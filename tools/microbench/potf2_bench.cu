// Micro-benchmark (development aid): cycle counts of the in-tile factorisation pieces, one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o potf2_bench potf2_bench.cu && ./potf2_bench
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../kadal_b200/csrc/kb200_kernels.cu"

using namespace kb200;

// ---- variant A: library rsqrt()/sqrt() path (the first version)
__device__ __noinline__ int potf2_32_A(double* B, double* dinv, int lane, int base) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = (c <= lane) ? B[lane * LDP + c] : 0.0;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    double piv = __shfl_sync(FULLMASK, a[k], k);
    if (!(piv > 0.0)) { if (!bad) bad = base + k + 1; piv = 1.0; }
    const double inv = rsqrt(piv);
    double dk = piv * inv;
    dk = fma(fma(-dk, dk, piv), 0.5 * inv, dk);
    const double l = a[k] * inv;
#pragma unroll
    for (int c = k + 1; c < 32; ++c) {
      const double lc = __shfl_sync(FULLMASK, l, c);
      a[c] = fma(-l, lc, a[c]);
    }
    a[k] = (lane == k) ? dk : l;
    if (lane == k) dinv[k] = inv;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) if (c <= lane) B[lane * LDP + c] = a[c];
  return bad;
}

// ---- variant C: shared-memory broadcast of the column (one STS + wide LDS instead of shuffles)
__device__ __noinline__ int potf2_32_C(double* B, double* dinv, double* lbuf /*[2][32]*/, int lane, int base) {
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
    const double l = a2 * h;
    double* lb = lbuf + (k & 1) * 32;
    lb[lane] = l;
    __syncwarp();
    if (k + 1 < 32) a[k + 1] = fma(-l, lb[k + 1], a[k + 1]);
#pragma unroll
    for (int c = k + 2; c < 32; ++c) a[c] = fma(-l, lb[c], a[c]);
    a[k] = (lane == k) ? dk : l;
    if (lane == k) dinv[k] = h + h;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) if (c <= lane) B[lane * LDP + c] = a[c];
  return bad;
}

__global__ void __launch_bounds__(512, 1) bench_kernel(const double* Ain, double* Lout, long long* cyc, int variant) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* S = reinterpret_cast<double*>(smem_raw);
  double* trow = S + TB * LDP + 8;
  double* dinv = trow + 8 * TLD;
  double* W8s = dinv + TB;
  double* lbuf = W8s + 1024;
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < TB * TB; i += blockDim.x) { int r = i / TB, c = i % TB; if (c <= r) S[r * LDP + c] = Ain[i]; }
  for (int i = tid; i < 8 * TLD; i += blockDim.x) trow[i] = 1.0;
  __syncthreads();
  long long t0 = clock64();
  if (variant == 0) {
    potf2_blocked(S, trow, dinv, W8s, lbuf, &s_bad);
  } else if (variant == 4) {
    potf2_simple(S, trow, dinv, W8s, &s_bad, 2);
  } else if (warp == 0) {
    // a single 32x32 block, 4 repetitions on the 4 diagonal blocks (un-updated: still SPD)
    for (int kb = 0; kb < 4; ++kb) {
      int o = 32 * kb;
      if (variant == 1) potf2_32_warp(S + o * LDP + o, dinv + o, lbuf, lane, o);
      else if (variant == 2) potf2_32_A(S + o * LDP + o, dinv + o, lane, o);
      else if (variant == 3) potf2_32_C(S + o * LDP + o, dinv + o, lbuf, lane, o);
      else if (variant == 5) { inv8_warp(S + o * LDP + o, dinv + o, W8s + kb * 256, lane); }
      else if (variant == 6) { strip_trsm32(S + (96) * LDP + o, LDP, S + o * LDP + o, W8s + kb * 256, lane); }
      else if (variant == 7) { trailing_update32(S, trow, 0, 3, 4, 0, 1, lane); }
      __syncwarp();
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  for (int i = tid; i < TB * TB; i += blockDim.x) { int r = i / TB, c = i % TB; Lout[i] = (c <= r) ? S[r * LDP + c] : 0.0; }
}

int main() {
  const int n = TB;
  std::vector<double> A(n * n), L(n * n);
  // SPD: A = M M^T / n + I
  std::vector<double> M(n * n);
  unsigned s = 12345;
  for (auto& v : M) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double t = 0;
      for (int k = 0; k < n; ++k) t += M[i * n + k] * M[j * n + k];
      A[i * n + j] = t / n + (i == j ? 1.0 : 0.0);
    }
  // reference Cholesky
  std::vector<double> R(A);
  for (int k = 0; k < n; ++k) {
    R[k * n + k] = sqrt(R[k * n + k]);
    for (int r = k + 1; r < n; ++r) R[r * n + k] /= R[k * n + k];
    for (int c = k + 1; c < n; ++c)
      for (int r = c; r < n; ++r) R[r * n + c] -= R[r * n + k] * R[c * n + k];
  }
  double *dA, *dL; long long* dc;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
  size_t smem = (size_t)(TB * LDP + 8 + 8 * TLD + TB + 1024 + 64) * 8;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"potf2_blocked (whole tile, 16 warps)", "4 x potf2_32_warp (current)", "4 x potf2_32_A (library rsqrt)",
                         "4 x potf2_32_C (smem broadcast)", "potf2_simple (whole tile)", "4 x inv8_warp", "4 x strip_trsm32", "4 x one trailing item"};
  for (int v = 0; v < 8; ++v) {
    long long best = 1LL << 60;
    for (int rep = 0; rep < 3; ++rep) {
      bench_kernel<<<1, 512, smem>>>(dA, dL, dc, v);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      if (c < best) best = c;
    }
    cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
    double err = 0;
    if (v == 0 || v == 4)
      for (int i = 0; i < n * n; ++i) err = fmax(err, fabs(L[i] - R[i]));
    printf("%-45s %8lld cycles   (max |L - Lref| = %.2e)\n", names[v], best, err);
  }
  return 0;
}
