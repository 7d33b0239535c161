// kadal_b200 — kernel argument blocks and launchers (internal; the public ABI is include/kb200.h)
#pragma once
#include <cuda_runtime.h>
#include "kb200_corr.cuh"

namespace kb200 {

enum { CORR_ASSEMBLE = 0, CORR_PREDICT = 1 };

struct CorrArgs {
  CorrSpec spec;
  int mode;
  int n;                 // training samples
  int nt;                // tiles per dimension (ceil(n / 128))
  const double* X;       // [n][d] training inputs (already normalised)
  const double* cand;    // [P][cand_stride]
  int cand_stride;
  // assembly
  double* out;           // tile storage (assembly: [P][tiles_per_mat] tiles; predict: chunk [pt][nt] or null)
  long tiles_per_mat;
  int add_eps;
  // prediction
  const double* xpts;    // [npts][d] raw points
  long npts;
  int normtype;          // 0 none, 1 'default' (lb,ub), 2 'std' (mean,std)
  const double* norm_a;
  const double* norm_b;
  const double* pscale;  // [d] prediction, single Gaussian kernel: per-dimension scale of the pre-scaled coordinates
                         //     (RN(1/theta_k); KPLS: sqrt(sum_c pls_kc^2 / theta_c^2)), or null (general path)
  const double* alpha;   // [nt*128] zero padded
  const double* wvec;    // [nt*128] zero padded, or null
  double* fdot;          // [npts]  psi^T alpha   (tj_split: [nt][part_stride] partial sums per sample tile)
  double* udot;          // [npts]  psi^T w
  int tj_split;          // prediction: blockIdx.y = sample tile, one tile pair per CTA (few point tiles: fills the machine)
  long part_stride;
  int skip_pad;          // prediction: do not store the padded 8-column blocks of the last sample tile (the half-SM variance
                         // kernels never read them, kb200_half.cu; the full-SM kernels do: leave 0 for those)
};

struct CholArgs {
  double* L;             // tile storage of the matrices being factorised (in place)
  double* W8;            // [P][nt][16][64] inverses of the 8x8 diagonal blocks of every diagonal tile
  long tiles_per_mat;
  int nt, npad, n;
  // fused forward substitution
  int nq;                // RHS columns (1 + p)
  const double* R;       // [nq][npad] shared RHS ([y F]^T, zero padded)
  double* Z;             // [P][nq][npad]
  double* logdet;        // [P]  sum ln L_ii
  int* info;             // [P]
  int naive;             // debug: plain-FMA GEMM instead of the DMMA pipeline
  int i0;                // half-SM panel: first tile row of this launch is j + 1 + i0
  int rl_mode;           // right-looking variant: 0 off, 1 solve-only (T), 2 trailing update (U), 3 U of the next tile column only
                         // (look-ahead), 4 prediction: U, and the CTAs of the next sample tile go on to solve it
  double* Rw;            // [P][nq][npad] working right-hand sides of the right-looking variant
  int fuse_diag;         // panel CTA (j+1, j) also finishes diagonal tile j+1 (no diag launches for j >= 1)
  // prediction-as-extra-rows
  int pred_mode;
  double* chunk;         // [npt_tiles][nt] tiles (psi in, L^-1 psi out)
  double* ssq;           // [npt_tiles*128]
};

struct FinArgs {
  const double* Z;
  const double* logdet;
  const int* info;
  const double* cand;
  int cand_stride;
  int nq, npad, n;
  int trainvar;
  double nsamp;
  double* nll;           // [P]
  double* sigma2;        // [P] or null
  double* beta;          // [P][p] or null
  double* btb;           // [P] or null  (B^T B = F^T Psi^-1 F, p = 1)
  double* resid;         // [P][npad] or null  (a - B beta)
};

struct BackArgs {
  const double* L;       // tiles of one matrix
  const double* W8;      // [nt][16][64] inverses of its 8x8 diagonal blocks
  int nt, npad, nq;
  const double* rhs;     // [nq][npad]
  double* out;           // [nq][npad]
  double* work;          // [nq][npad] scratch: the right-hand sides as the solve consumes them
};
inline int trsv_launches(int nt) { return 2 * nt - 1; }

struct DecodeArgs {
  const double* x;       // [P][nbhyp]
  long P;
  int nbhyp, ntheta, nkernel;
  int iso;               // iso_gaussian: one theta replicated
  int theta_is_linear;   // x already holds theta (not log10) — unused by the public API
  int trainvar;
  int layout;            // bit0: tuned nugget in x, bit1: kernel weights in x
  double eps_fixed;      // 10**KrigInfo['nugget'] when the nugget is fixed
  const double* wgkf_given;  // device [nkernel] or null
  double* cand;
  int cand_stride;
};

struct EpiArgs {
  long npts;
  double beta, sigma2, btb, ymin, limit;
  int limit_ge, need_var;
  const double* fdot;
  const double* udot;
  int nparts;            // > 1: fdot / udot are nparts partial sums, part_stride apart, added in order
  long part_stride;
  const double* ssq;
  double *pred, *ssqr, *s, *ei, *poi, *pof, *ufun;
  unsigned long long* stats;   // [0] = #(SSqr == 0), [1] = any nonzero EI term
};

struct PredSmallArgs {
  int n, d, nt, ntheta;
  const double* X;       // [n][d] training inputs
  const double* cand;    // parameter row of the fitted model
  const double* pscale;  // [d] per-dimension scale of the pre-scaled coordinates (see CorrArgs)
  const double* L;       // predictor tiles L_00, L_10, L_11
  const double* W8;      // [nt][1024]
  const double* alpha;   // [nt*128]
  const double* wvec;    // [nt*128]
  const double* xpts;    // [npts][d] raw points
  long npts;
  int normtype;
  const double* norm_a;
  const double* norm_b;
  EpiArgs epi;           // outputs / scalars (fdot, udot, ssq unused)
};

// KPLS assembly from cached projected-distance planes (kb200_kpls.cu)
struct KplsAsmArgs {
  int n, h, nkernel, add_eps;
  int kid[MAXK];
  long tiles_per_mat;
  const double* planes_sq;   // [h][tiles_per_mat] tiles: sum_k (dx_k)^2 pls_kc^2   (gaussian) or null
  const double* planes_abs;  // [h][tiles_per_mat] tiles: sum_k |dx_k| pls_kc       (exponential) or null
  const double* cand;        // [P][cand_stride]
  int cand_stride;
  double* out;               // [P][tiles_per_mat] tiles
};
cudaError_t launch_kpls_planes(const double* X, int n, int d, int h, const double* coef, int absmode, double* planes,
                               long tiles_per_mat, cudaStream_t st);
cudaError_t launch_kpls_assemble(const KplsAsmArgs& a, int P, cudaStream_t st);

// 2-objective EHVI (kb200_ehvi.cu)
struct EhviArgs {
  int k;                       // Pareto points
  double r0, r1;               // reference point
  const double* S;             // [k][2] front sorted on objective 1
  const double* c2;            // [k] objective-2 values ascending
  const double* splus;         // [(k+1)][(k+1)] dominated hyper-volume per cell
  const double *mu0, *mu1, *s0, *s1;   // candidate i at index i * inc
  long inc;
  double sign;                 // +1: exi2d, -1: ehvicalc_vec's -exi2d
  double* out;                 // [npop]
};
cudaError_t launch_ehvi_tables(const double* P, int k, double r0, double r1, double* S, double* c2, double* splus,
                               cudaStream_t st);
cudaError_t launch_ehvi_score(const EhviArgs& a, long npop, cudaStream_t st);

// 3-objective EHVI (kb200_ehvi3d.cu); maximisation convention of the reference's KMAC extension
struct Ehvi3Args {
  int n;                       // front points (mutually non-dominated)
  long npop;                   // candidates
  int mode;                    // 0: reference semantics (order-dependent `break`), 1: every candidate on its own
  double r0, r1, r2;           // reference point (lower corner)
  const double* front;         // [n][3]
  int* rank;                   // [3][n] rank of point i per axis
  double* coord;               // [3][n] sorted coordinates per axis
  int* hd;                     // [n][n] highest dominator of node (x, y)
  double *xlim, *ylim;         // [n][n]
  double *slice, *chunk;       // [n + 1][n][n]
  const double *mean, *sdev;   // candidate i, objective c at [i * inc + c * cinc]
  long inc, cinc;
  double *E, *G, *EX;          // [3][n + 1][npop]
  long* first;                 // [3][n + 1]
  int ysplit;                  // CTAs per z level along y (ehvi3_ysplit)
  double* partial;             // [(n + 1) * ysplit][npop]
  double sign;
  double* out;                 // [npop]
};
cudaError_t launch_ehvi3(const Ehvi3Args& a, cudaStream_t st, int phase = 3);
int ehvi3_ysplit(int n, long npop);

// Saltelli sums (kb200_sobol.cu)
constexpr int SOBOL_BLOCKS = 296;       // 2 x 148
cudaError_t launch_sobol_compose(const double* A, const double* B, const unsigned char* from_a, long n, int d, double* C,
                                 cudaStream_t st);
cudaError_t launch_sobol_points(const unsigned long long* dirnum, int bits, int dims, long long i0, long long n,
                                const double* lo, const double* hi, double* outA, double* outB, int split, cudaStream_t st);
cudaError_t launch_sobol_dot(const double* u, const double* v, const double* w, long n, int mode, double* partial,
                             double* acc0, double* acc1, cudaStream_t st);

cudaError_t launch_corr(const CorrArgs& a, int grid_x, int grid_y, cudaStream_t st);
constexpr int AK_BLOCKS = 592;          // 4 x 148
constexpr int AK_PARTIAL_BYTES = 40;    // {double min_u; long long arg; unsigned long long c0, cm, cp;}
cudaError_t launch_akmcs_reduce(const double* G, const double* S, long n, long base, void* out, int nblocks, cudaStream_t st);
cudaError_t launch_identity_chunk(double* chunk, int tiles, int nt, long pt0, cudaStream_t st);
// one-sample append to a factor (believer / sequential refits)
cudaError_t launch_append_extract(const double* tiles, int n_prev, int nt_prev, double* vec, double* dnn,
                                  cudaStream_t st);
cudaError_t launch_append_row(double* tiles, const double* v, int n_prev, int nt_prev, int nq, int npad_prev,
                              int npad, const double* Zprev, const double* R, double* Z, const double* logdet_prev,
                              double* logdet, int* info, const double* dnn, cudaStream_t st);
cudaError_t launch_loo(const double* y, const double* alpha, const double* w, const double* ssq, double btb, int n,
                       long p0, int np, double* out, cudaStream_t st, double fast_sigma2 = 0.0, double beta = 0.0);
cudaError_t launch_predict_small(const PredSmallArgs& a, int sm_count, cudaStream_t st);
// per-dimension coordinate scales of a single-Gaussian-kernel predictor (pls2 = squared KPLS coefficients [d][h] or null)
cudaError_t launch_pred_scale(const double* cand_row, const double* pls2, int d, int ntheta, int kpls, double* out, cudaStream_t st);
bool predict_small_supported(int n, int d, int nkernel, int kid0, int kpls);
cudaError_t launch_chol_diag(const CholArgs& a, int j, int nmat, cudaStream_t st);
cudaError_t launch_chol_panel(const CholArgs& a, int j, int grid_x, int grid_y, cudaStream_t st);
// half-SM variants (kb200_half.cu): two CTAs per SM
cudaError_t launch_chol_diag_half(const CholArgs& a, int j, int nmat, cudaStream_t st);
cudaError_t launch_chol_panel_half(const CholArgs& a, int j, int grid_x, int grid_y, cudaStream_t st);
// small n: assembly + factorisation + forward substitution of one candidate per CTA in shared memory (kb200_half.cu)
int lik_small_pad(int n, int d);
bool half_kernels_skip_padding();   // the half-SM kernels skip the identity padding of a ragged last tile (KB200_RAGGED)
cudaError_t launch_lik_small(const CorrArgs& c, const CholArgs& a, int nmat, cudaStream_t st);
cudaError_t launch_rl_rhs_update(const CholArgs& a, int k, int nmat, cudaStream_t st, int ntiles = -1);
cudaError_t launch_w8_from_tiles(const CholArgs& a, int nmat, cudaStream_t st);
cudaError_t launch_finalize(const FinArgs& a, int nmat, cudaStream_t st);
cudaError_t launch_backsolve(const BackArgs& a, cudaStream_t st);
cudaError_t launch_import_U(const double* U, int n, int nt, double* tiles, cudaStream_t st);
cudaError_t launch_export(const double* tiles, int n, double* out, int what, cudaStream_t st);
cudaError_t launch_add_diag(double* tiles, int n, const double* eps_ptr, cudaStream_t st);
cudaError_t launch_forwardsolve(const BackArgs& a, cudaStream_t st);
cudaError_t launch_decode(const DecodeArgs& a, cudaStream_t st);
cudaError_t launch_epilogue(const EpiArgs& a, cudaStream_t st);
cudaError_t launch_debug_exp(const double* x, long n, int mode, double* y, cudaStream_t st);
cudaError_t launch_fp64_probe(int blocks, int iters, int mode, double* sink, cudaStream_t st);

}  // namespace kb200
