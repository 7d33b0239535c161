// kadal_b200 — 3-objective expected hyper-volume improvement on the device (second half).
//
// Replaces kmac.ehvi3d_sliceupdate_multi, the reference's C++ extension
// (kadal/extern/HDYE_3D_Update/ehvi_multi.cc:113-341 behind kmac_wrapper.cpp:73-118), as called per
// population by optim_tools/ehvi/EHVIcomputation.py:235.  Maximisation convention; the grid spanned by the
// sorted front coordinates (plus the reference point below and +inf above) has (n+1)^3 cells, and a
// candidate's EHVI is the sum over the non-dominated cells of
//     psi1 psi2 psi3 - S^- G1 G2 G3 - sum_d slice_d (..) - sum_d v_d (..)          (ehvi_multi.cc:321-325)
// where everything that depends on the front alone (S^- "chunk", the three slices, the three one-dimensional
// limits) is shared by all candidates and everything that depends on the candidate factorises per axis.
// The reference walks cells x candidates on one core, caching the per-axis terms in 12 static 1001 x 1001
// arrays.  Here:
//   ehvi3_rank_kernel    ranks of the front points per axis (stable), sorted coordinates
//   ehvi3_node_kernel    per grid node: highest dominator, x- and y-limits             (:181-196, O(n) each)
//   ehvi3_slice_kernel   per z level: the dominated area below every node              (:217-236)
//   ehvi3_chunk_kernel   per node: running dominated volume over the z levels          (:212-215)
//   ehvi3_axis_kernel    per axis, cell and candidate: psi, G, psi - G (cl - r)        (:294-320)
//   ehvi3_first_kernel   per axis and cell: first candidate whose psi < 1e-15 -- the reference leaves its
//                        candidate loop with `break` there, so LATER candidates lose the cell; reproduced
//                        (mode 0) or not (mode 1: every candidate as if evaluated alone)
//   ehvi3_score_kernel   grid = candidate blocks x z levels, thread = candidate: the cell constants of one
//                        grid row are staged in shared memory once per CTA, the candidate's axis terms are
//                        coalesced loads; partial sums per z level
//   ehvi3_sum_kernel     sum over the z levels in order, sign
// Every product / sum is rounded separately in the reference's order and with the reference's truncated
// constants (ehvi_consts.h:9-10), so a term differs from the reference only through the last ulp of erf/exp.
#include <cmath>
#include "kb200_common.cuh"
#include "kb200_kernels.h"

namespace kb200 {
namespace {

constexpr double E3_SQRT_TWO = 1.41421356237;        // ehvi_consts.h:9 (truncated in the reference)
constexpr double E3_INV_SQRT_2PI = 0.3989422804;     // ehvi_consts.h:10
constexpr double E3_THRESH = 0.000000000000001;      // ehvi_multi.cc:298,303,308,311
constexpr int E3_THREADS = 128;

__device__ __forceinline__ double e3_pdf(double x) {          // helper.h:76-78
  return __dmul_rn(E3_INV_SQRT_2PI, exp(-(__dmul_rn(x, x) / 2.0)));
}
__device__ __forceinline__ double e3_cdf(double x) {          // helper.h:81-83
  return __dmul_rn(0.5, __dadd_rn(1.0, erf(__ddiv_rn(x, E3_SQRT_TWO))));
}
__device__ __forceinline__ double e3_psi(double fmax, double corner, double mu, double s) {   // helper.h:86-88
  const double z = __ddiv_rn(__dsub_rn(corner, mu), s);
  return __dadd_rn(__dmul_rn(s, e3_pdf(z)), __dmul_rn(__dsub_rn(fmax, mu), e3_cdf(z)));
}

__global__ void ehvi3_rank_kernel(Ehvi3Args A) {
  const int n = A.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double f0 = A.front[3 * i], f1 = A.front[3 * i + 1], f2 = A.front[3 * i + 2];
    int r0 = 0, r1 = 0, r2 = 0;
    for (int j = 0; j < n; ++j) {
      const double g0 = A.front[3 * j], g1 = A.front[3 * j + 1], g2 = A.front[3 * j + 2];
      r0 += (g0 < f0) || (g0 == f0 && j < i);
      r1 += (g1 < f1) || (g1 == f1 && j < i);
      r2 += (g2 < f2) || (g2 == f2 && j < i);
    }
    A.rank[i] = r0; A.rank[n + i] = r1; A.rank[2 * n + i] = r2;
    A.coord[r0] = f0; A.coord[n + r1] = f1; A.coord[2 * n + r2] = f2;
  }
}

// node (a, b): hd[a + b n]   = highest z rank among the points with x rank >= a and y rank >= b (-1: none)
//              xlim[a + b n] = largest f0 - r0 among the points with y rank >= a and z rank >= b (0: none)
//              ylim[a + b n] = largest f1 - r1 among the points with x rank >= a and z rank >= b (0: none)
__global__ void ehvi3_node_kernel(Ehvi3Args A) {
  const int n = A.n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    const int a = idx % n, b = idx / n;
    int hd = -1;
    double mx = -INFINITY, my = -INFINITY;
    for (int i = 0; i < n; ++i) {
      const int xr = A.rank[i], yr = A.rank[n + i], zr = A.rank[2 * n + i];
      if (xr >= a && yr >= b) hd = max(hd, zr);
      if (yr >= a && zr >= b) mx = fmax(mx, A.front[3 * i]);
      if (xr >= a && zr >= b) my = fmax(my, A.front[3 * i + 1]);
    }
    A.hd[idx] = hd;
    A.xlim[idx] = mx == -INFINITY ? 0.0 : __dsub_rn(mx, A.r0);
    A.ylim[idx] = my == -INFINITY ? 0.0 : __dsub_rn(my, A.r1);
  }
}

// one thread per z level: slice[z][x + y n] = dominated area of the level-z cross-section below node (x, y)
__global__ void ehvi3_slice_kernel(Ehvi3Args A) {
  const int n = A.n;
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z > n) return;
  double* sl = A.slice + (size_t)z * n * n;
  const double* __restrict__ X = A.coord;
  const double* __restrict__ Y = A.coord + n;
  const int* __restrict__ hd = A.hd;
  for (int y = 0; y < n; ++y) {
    const double wy = __dsub_rn(Y[y], A.r1);
    // rows do not overlap: `restrict` lets the loads of the row above run ahead of the stores of this row
    // (the chain through `left` is a few cycles per node, an un-hoisted load is an L2 round trip per node)
    const double* __restrict__ prev = sl + (size_t)(y > 0 ? y - 1 : 0) * n;
    double* __restrict__ cur = sl + (size_t)y * n;
    const int* __restrict__ hrow = hd + (size_t)y * n;
    double left = 0.0, upleft = 0.0;
    for (int x0 = 0; x0 < n; x0 += 8) {
      // eight nodes per trip: their loads are issued together (one L2 round trip per eight nodes instead of
      // one per node: the chain through `left` is only a few cycles long)
      double upv[8], xv[8];
      int hv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int x = min(x0 + u, n - 1);
        upv[u] = y > 0 ? prev[x] : 0.0;
        hv[u] = hrow[x];
        xv[u] = X[x];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int x = x0 + u;
        if (x < n) {
          const double up = upv[u];
          double v;
          if (hv[u] < z) {
            if (x > 0 && y > 0) v = __dadd_rn(__dsub_rn(up, upleft), left);
            else if (y > 0) v = up;
            else if (x > 0) v = left;
            else v = 0.0;
          } else {
            v = __dmul_rn(__dsub_rn(xv[u], A.r0), wy);
          }
          cur[x] = v;
          left = v;
          upleft = up;
        }
      }
    }
  }
}

// one thread per node: chunk[z] = chunk[z - 1] + slice[z - 1] * (length of z cell z - 1)
__global__ void ehvi3_chunk_kernel(Ehvi3Args A) {
  const int n = A.n;
  const double* Z = A.coord + 2 * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    double c = 0.0;
    A.chunk[idx] = 0.0;
    for (int z = 1; z <= n; ++z) {
      const double len = __dsub_rn(Z[z - 1], z == 1 ? A.r2 : Z[z - 2]);
      c = __dadd_rn(c, __dmul_rn(A.slice[(size_t)(z - 1) * n * n + idx], len));
      A.chunk[(size_t)z * n * n + idx] = c;
    }
  }
}

// E / G / EX [axis][cell][candidate]
__global__ void ehvi3_axis_kernel(Ehvi3Args A) {
  const int n = A.n, n1 = n + 1;
  const long total = 3L * n1 * A.npop;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long i = idx % A.npop;
    const int cell = (int)((idx / A.npop) % n1), c = (int)(idx / (A.npop * n1));
    const double r = c == 0 ? A.r0 : (c == 1 ? A.r1 : A.r2);
    const double* co = A.coord + c * n;
    const double cl = cell == 0 ? r : co[cell - 1];
    const double cu = cell == n ? INFINITY : co[cell];
    const double mu = A.mean[i * A.inc + c * A.cinc], s = A.sdev[i * A.inc + c * A.cinc];
    const double e = __dsub_rn(e3_psi(r, cl, mu, s), e3_psi(r, cu, mu, s));
    const double g = __dsub_rn(e3_cdf(__ddiv_rn(__dsub_rn(cu, mu), s)), e3_cdf(__ddiv_rn(__dsub_rn(cl, mu), s)));
    A.E[idx] = e;
    A.G[idx] = g;
    A.EX[idx] = __dsub_rn(e, __dmul_rn(g, __dsub_rn(cl, r)));
  }
}

// first[axis][cell] = first candidate with E < 1e-15 (npop if none); one CTA per (axis, cell)
__global__ void __launch_bounds__(E3_THREADS)
ehvi3_first_kernel(Ehvi3Args A) {
  __shared__ long best[E3_THREADS];
  const double* e = A.E + (size_t)blockIdx.x * A.npop;
  long f = A.npop;
  if (A.mode == 0)
    for (long i = threadIdx.x; i < A.npop; i += E3_THREADS)
      if (e[i] < E3_THRESH) { f = i; break; }
  best[threadIdx.x] = f;
  __syncthreads();
  for (int o = E3_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) best[threadIdx.x] = min(best[threadIdx.x], best[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) A.first[blockIdx.x] = best[0];
}

// smem: per x of the current grid row: skip flag, S^-, slice_x, slice_y, slice_z, v_x, v_y, v_z, first_x
__global__ void __launch_bounds__(E3_THREADS)
ehvi3_score_kernel(Ehvi3Args A) {
  extern __shared__ double sm[];
  const int n = A.n, n1 = n + 1, z = blockIdx.y;
  const int ys = blockIdx.z, nys = gridDim.z;            // this CTA's share of the grid rows of level z
  const int y_lo = (int)((long)n1 * ys / nys), y_hi = (int)((long)n1 * (ys + 1) / nys);
  double* c_sm = sm;             // [n1]
  double* c_s0 = c_sm + n1;
  double* c_s1 = c_s0 + n1;
  double* c_s2 = c_s1 + n1;
  double* c_v0 = c_s2 + n1;
  double* c_v1 = c_v0 + n1;
  double* c_v2 = c_v1 + n1;
  long* c_first = reinterpret_cast<long*>(c_v2 + n1);           // < 0: cell skipped
  const long i = (long)blockIdx.x * E3_THREADS + threadIdx.x;
  const bool mine = i < A.npop;
  const size_t np = (size_t)A.npop;
  const double *X = A.coord, *Y = A.coord + n, *Z = A.coord + 2 * n;
  const double cl2 = z == 0 ? A.r2 : Z[z - 1], cu2 = z == n ? INFINITY : Z[z];
  const double len2 = __dsub_rn(cu2, cl2);
  const double* chunk = A.chunk + (size_t)z * n * n;
  const double* slice = A.slice + (size_t)z * n * n;
  const long fz = A.first[2 * n1 + z];
  double e3 = 0.0, g3 = 0.0, ex3 = 0.0;
  if (mine) {
    const size_t o = ((size_t)2 * n1 + z) * np + i;
    e3 = A.E[o]; g3 = A.G[o]; ex3 = A.EX[o];
  }
  double acc = 0.0;
  for (int y = y_lo; y < y_hi; ++y) {
    const double cl1 = y == 0 ? A.r1 : Y[y - 1], cu1 = y == n ? INFINITY : Y[y];
    const double len1 = __dsub_rn(cu1, cl1);
    const long fyz = min(fz, A.first[n1 + y]);
    // reference semantics: an early exit in front of every candidate of this CTA empties the whole grid row
    if ((long)blockIdx.x * E3_THREADS >= fyz) continue;        // (uniform: depends on the CTA, y and z only)
    __syncthreads();
    for (int x = threadIdx.x; x <= n; x += E3_THREADS) {        // ehvi_multi.cc:241-282
      const double cl0 = x == 0 ? A.r0 : X[x - 1], cu0 = x == n ? INFINITY : X[x];
      const double len0 = __dsub_rn(cu0, cl0);
      const bool skip = len0 == 0.0 || len1 == 0.0 || len2 == 0.0 || (x < n && y < n && A.hd[x + y * n] >= z);
      double sminus = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0;
      if (!skip) {
        if (x > 0 && y > 0) {
          sminus = chunk[(x - 1) + (y - 1) * n];
          s0 = x == n ? 0.0 : __ddiv_rn(__dsub_rn(chunk[x + (y - 1) * n], sminus), len0);
          s1 = y == n ? 0.0 : __ddiv_rn(__dsub_rn(chunk[(x - 1) + y * n], sminus), len1);
          s2 = slice[(x - 1) + (y - 1) * n];
        } else {
          s0 = (y == 0 || x == n) ? 0.0 : __ddiv_rn(chunk[x + (y - 1) * n], len0);
          s1 = (x == 0 || y == n) ? 0.0 : __ddiv_rn(chunk[(x - 1) + y * n], len1);
        }
      }
      c_sm[x] = sminus; c_s0[x] = s0; c_s1[x] = s1; c_s2[x] = s2;
      c_v0[x] = (y == n || z == n) ? 0.0 : A.xlim[y + z * n];
      c_v1[x] = (x == n || z == n) ? 0.0 : A.ylim[x + z * n];
      double v2 = 0.0;
      if (x < n && y < n) {
        const int h = A.hd[x + y * n];
        v2 = h < 0 ? 0.0 : __dsub_rn(Z[h], A.r2);
      }
      c_v2[x] = v2;
      c_first[x] = skip ? -1 : min(fyz, A.first[x]);
    }
    __syncthreads();
    if (!mine) continue;
    const size_t oy = ((size_t)n1 + y) * np + i;
    const double e2 = A.E[oy], g2 = A.G[oy], ex2 = A.EX[oy];
    const bool dead2 = A.mode != 0 && (e2 < E3_THRESH || e3 < E3_THRESH);
    for (int x = 0; x <= n; ++x) {
      if (i >= c_first[x]) continue;                            // skipped cell, or a `break` in front of us
      const size_t ox = (size_t)x * np + i;
      const double e1 = A.E[ox];
      if (A.mode != 0 && (dead2 || e1 < E3_THRESH)) continue;
      const double psiprod = __dmul_rn(__dmul_rn(e1, e2), e3);
      if (!(psiprod > E3_THRESH)) continue;
      const double g1 = A.G[ox], ex1 = A.EX[ox];
      double t = __dsub_rn(psiprod, __dmul_rn(__dmul_rn(__dmul_rn(c_sm[x], g1), g2), g3));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_s0[x], g2), g3), ex1));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_v0[x], ex2), ex3), g1));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_s1[x], g1), g3), ex2));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_v1[x], ex1), ex3), g2));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_s2[x], g1), g2), ex3));
      t = __dsub_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(c_v2[x], ex1), ex2), g3));
      if (t > 0.0) acc = __dadd_rn(acc, t);
    }
  }
  if (mine) A.partial[((size_t)z * nys + ys) * np + i] = acc;
}

__global__ void ehvi3_sum_kernel(Ehvi3Args A) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < A.npop; i += (long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < (A.n + 1) * A.ysplit; ++q) s = __dadd_rn(s, A.partial[(size_t)q * A.npop + i]);
    A.out[i] = A.sign * s;
  }
}

}  // namespace

size_t ehvi3_smem_bytes(int n) { return (size_t)(n + 1) * 8 * sizeof(double); }

// grid rows of a z level are split over `ysplit` CTAs so that a small population still fills the machine
// (about eight 128-thread CTAs per SM); every split re-reads the level's tables from L2, so no more than that
int ehvi3_ysplit(int n, long npop) {
  const long ctas = ((npop + E3_THREADS - 1) / E3_THREADS) * (n + 1);
  long want = (148L * 8 + ctas - 1) / ctas;
  return (int)max(1L, min(want, min((long)(n + 1), 16L)));
}

// phase bit 0: front geometry (ranks, nodes, slices, chunks -- independent of the candidates); bit 1: score a.npop candidates
cudaError_t launch_ehvi3(const Ehvi3Args& a, cudaStream_t st, int phase) {
  const int n = a.n, n1 = n + 1;
  auto blocks = [](long work, int per) { return (unsigned)max(1L, min((work + per - 1) / per, 148L * 16)); };
#define E3_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return e__; } while (0)
  if (n > 0 && (phase & 1)) {
    ehvi3_rank_kernel<<<blocks(n, 128), 128, 0, st>>>(a);
    E3_CHECK();
    ehvi3_node_kernel<<<blocks((long)n * n, 128), 128, 0, st>>>(a);
    E3_CHECK();
    ehvi3_slice_kernel<<<(n1 + 31) / 32, 32, 0, st>>>(a);
    E3_CHECK();
    ehvi3_chunk_kernel<<<blocks((long)n * n, 128), 128, 0, st>>>(a);
    E3_CHECK();
  }
  if (a.npop <= 0 || !(phase & 2)) return cudaSuccess;
  ehvi3_axis_kernel<<<blocks(3L * n1 * a.npop, 256), 256, 0, st>>>(a);
  E3_CHECK();
  ehvi3_first_kernel<<<3 * n1, E3_THREADS, 0, st>>>(a);
  E3_CHECK();
  const size_t sm = ehvi3_smem_bytes(n);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ehvi3_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  ehvi3_score_kernel<<<dim3((unsigned)((a.npop + E3_THREADS - 1) / E3_THREADS), n1, a.ysplit), E3_THREADS, sm, st>>>(a);
  E3_CHECK();
  ehvi3_sum_kernel<<<blocks(a.npop, 128), 128, 0, st>>>(a);
  E3_CHECK();
#undef E3_CHECK
  return cudaSuccess;
}

}  // namespace kb200
