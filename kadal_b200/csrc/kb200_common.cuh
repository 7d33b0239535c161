// kadal_b200 — common device helpers (sm_100a only).
//
// Data layout used by every kernel ("tile-major, slab-swizzled"):
//   a symmetric/triangular n×n fp64 matrix is padded to NT×NT tiles of TB×TB (TB=128);
//   only the lower-triangular tiles (ti >= tj) are stored, tile (ti,tj) at index
//   ti*(ti+1)/2 + tj, each tile TB*TB doubles (128 KiB) contiguous in HBM.
//   Inside a tile the 128 columns are cut into 8 "slabs" of 16 columns; a slab is
//   128 rows × 16 doubles (one 128-byte line per row, 16 KiB per slab, contiguous), and
//   the four 32-byte chunks of a line are XOR-swizzled with (row & 3).
//   -> a slab is one cp.async.bulk (TMA 1-D) copy into shared memory, and the
//      mma.sync m8n8k4 fp64 fragment loads (8 rows × 4 consecutive doubles per
//      k-step) are bank-conflict free without padding.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kb200 {

constexpr int TB = 128;             // tile edge
constexpr int KS = 16;              // slab width (columns)
constexpr int NSLAB = TB / KS;      // 8 slabs per tile
constexpr int SLAB_ELEMS = TB * KS; // 2048 doubles = 16 KiB
constexpr int TILE_ELEMS = TB * TB; // 16384 doubles = 128 KiB
constexpr uint32_t SLAB_BYTES = SLAB_ELEMS * 8;

// kernel ids (KrigInfo['kernel'] strings are mapped by the host)
enum { K_GAUSS = 0, K_EXPO = 1, K_MAT32 = 2, K_MAT52 = 3 };

__host__ __device__ __forceinline__ long tri_index(int ti, int tj) {
  return (long)ti * (ti + 1) / 2 + tj;
}
// offset of element (r,c) inside a tile
__host__ __device__ __forceinline__ int tile_off(int r, int c) {
  return (c >> 4) * SLAB_ELEMS + r * KS + ((((c >> 2) & 3) ^ (r & 3)) << 2) + (c & 3);
}
// offset of element (r, k) inside one slab (k in [0,16))
__host__ __device__ __forceinline__ int slab_off(int r, int k) {
  return r * KS + (((k >> 2) ^ (r & 3)) << 2) + (k & 3);
}

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completes on mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// make generic-proxy writes visible to the async proxy (bulk copies issued afterwards)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;\n" ::: "memory");
}

// fp64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col).  SASS: DMMA.8x8x4
//   a : A[lane/4][lane%4]      b : B[lane%4][lane/4]
//   c0,c1 : C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// numpy's pairwise summation order for a contiguous reduction of length n <= 128
// (numpy/core/src/umath/loops_utils.h.src: pairwise_sum; n < 8: plain loop,
// else 8 strided partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the
// tail added in order).  `term(k)` yields the k-th summand.  (For n > 128 numpy recurses
// on halves; such reduction lengths are rejected by the host API.)
template <typename F>
__device__ __forceinline__ double np_sum_block(int lo, int n, F term) {
  if (n < 8) {
    double res = 0.0;
    for (int k = 0; k < n; ++k) res += term(lo + k);
    return res;
  }
  double r0 = term(lo + 0), r1 = term(lo + 1), r2 = term(lo + 2), r3 = term(lo + 3);
  double r4 = term(lo + 4), r5 = term(lo + 5), r6 = term(lo + 6), r7 = term(lo + 7);
  int nfull = n - (n & 7);
  for (int i = 8; i < nfull; i += 8) {
    r0 += term(lo + i + 0); r1 += term(lo + i + 1); r2 += term(lo + i + 2); r3 += term(lo + i + 3);
    r4 += term(lo + i + 4); r5 += term(lo + i + 5); r6 += term(lo + i + 6); r7 += term(lo + i + 7);
  }
  double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (int i = nfull; i < n; ++i) res += term(lo + i);
  return res;
}
// IEEE-correct a/b from a correctly rounded reciprocal rb = RN(1/b) (Markstein):
// q0 = a*rb is within 1 ulp, the fma residual correction rounds it correctly
// (b, a finite, no over/underflow in the quotient — true for all distances here).
__device__ __forceinline__ double div_by_rcp(double a, double b, double rb) {
  double q0 = a * rb;
  double r = fma(-b, q0, a);
  return fma(r, rb, q0);
}

}  // namespace kb200
