/* kb200.h — C ABI of kadal_b200: the B200-native drop-in for KADAL's Kriging hot path.
 *
 * KADAL (flowdiagnosticsitb/KADAL) has no FFI on this path: the boundary is two Python
 * callables and the KrigInfo dict.  These entry points are what a ctypes /
 * cffi / pybind binding of that path binds; each one names the reference interface it
 * replaces (paths relative to the KADAL repository root).
 *
 * Conventions
 *   - all matrices are C-contiguous row-major fp64; the caller owns every buffer;
 *   - the library owns only the opaque kb200_model (device copies + workspaces);
 *   - functions return 0 on success, <0 on a CUDA / argument error (kb200_last_error());
 *   - numerical failure is reported per item in info[] (0 ok, k>0 = first non-positive
 *     pivot at row k-1), never as a crash;
 *   - `*_dev` variants take device pointers and a cudaStream_t (as void*) and do no
 *     host<->device copies; the plain variants take host pointers and block.
 *   - stream ordering of the `*_dev` variants: with a non-NULL stream the call is asynchronous and
 *     ordered on that stream (inputs must be ready on it, outputs are ready on it).  With stream ==
 *     NULL the work runs on the model's own stream and the call BLOCKS until the outputs are complete.
 *   - a model supports ONE call in flight: its scratch buffers (decoded candidates, prediction slots,
 *     workspaces) are shared by every entry point.  The library enforces this itself -- each call first
 *     makes its stream wait for the last work of the previous call on the same model, whatever stream
 *     that ran on -- so calls on different streams are serialised, never concurrent.  Use one model per
 *     stream for concurrency.  A model must not be used from two host threads at once.
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef KB200_H
#define KB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb200_model kb200_model;

/* kernel ids = KrigInfo['kernel'] entries (kernelfunc.py:22-31) */
#define KB200_KERNEL_GAUSSIAN 0      /* also 'iso_gaussian' with KB200_FLAG_ISO */
#define KB200_KERNEL_EXPONENTIAL 1
#define KB200_KERNEL_MATERN32 2
#define KB200_KERNEL_MATERN52 3

#define KB200_FLAG_ISO 1             /* KrigInfo['kernel'] == ['iso_gaussian'] (likelihood_func.py:60-61) */
#define KB200_FLAG_NAIVE 2           /* debug: plain-FMA GEMM kernels instead of the DMMA pipeline */

/* outputs of kb200_predict (bit mask `want`), prediction.py:258-314 + akmcs.py:281-285 */
#define KB200_OUT_PRED 1
#define KB200_OUT_SSQR 2
#define KB200_OUT_S 4
#define KB200_OUT_EI 8
#define KB200_OUT_POI 16
#define KB200_OUT_POF 32
#define KB200_OUT_UFUN 64

#define KB200_NORM_NONE 0
#define KB200_NORM_DEFAULT 1         /* ((x-lb)/(ub-lb)-0.5)*2  samplingplan.py:77-83 */
#define KB200_NORM_STD 2             /* (x-mean)/std            prediction.py:164-165 */

int kb200_version(void);
const char* kb200_last_error(void);
int kb200_device_count(void);

/* Pinned host memory for the host-pointer entry points (optional; pageable works too). */
void* kb200_host_alloc(unsigned long long bytes);
void kb200_host_free(void* p);

/* Device-resident model = the part of KrigInfo that likelihood()/prediction() read:
 * X_norm|X (n x d), y|y_norm (n), F (n x p), plscoeff (d x h or NULL), kernel list.
 * Replaces the unpacking in likelihood_func.py:38-68 and prediction.py:122-157.
 * nsamp = KrigInfo['nsamp'] (used by the trainvar branch, likelihood_func.py:197). */
int kb200_model_create(kb200_model** out, int device, int n, int d, int p, int h,
                       const double* X, const double* y, const double* F, const double* pls,
                       const int* kernel_ids, int nkernel, double nsamp, int flags);
void kb200_model_destroy(kb200_model* m);

/* Batched negative concentrated ln-likelihood: one call evaluates
 *   likelihood(x[i], KrigInfo, 'default', trainvar)   for i < P
 * (likelihood_func.py:6-118 incl. nuggetset :121-165 and maincalc :168-213).
 * x is P x nbhyp with the reference's layout [log10 theta | log10 sigma2 | log10 nugget |
 * kernel weights]; nugget_log10 = KrigInfo['nugget'] when it is not tuned.
 * Non-PD candidates get nll = +inf and info != 0 (the reference raises LinAlgError). */
int kb200_lik_batch(kb200_model* m, const double* x, int P, int nbhyp, int trainvar,
                    double nugget_log10, double* nll, int* info);
int kb200_lik_batch_dev(kb200_model* m, const double* d_x, int P, int nbhyp, int trainvar,
                        double nugget_log10, double* d_nll, int* d_info, void* stream);

/* likelihood(x, KrigInfo, mode='all'): the quantities the reference writes back into
 * KrigInfo (likelihood_func.py:107-111): U (n x n, upper, with nugget), Psi (n x n, without
 * nugget), BE (p), SigmaSqr, NegLnLike; plus the decoded Theta (ntheta, log10), nugget
 * (log10) and wgkf (nkernel) of :82-84.  Any output pointer may be NULL.
 * Also installs the predictor state of this fit in the model. */
int kb200_fit_state(kb200_model* m, const double* x, int nbhyp, int trainvar,
                    double nugget_log10, double* U, double* Psi, double* BE, double* sigma2,
                    double* nll, int* info, double* theta_log10, double* nugget_out,
                    double* wgkf);

/* The same call after ONE sample was stacked under X, y and F -- what the Kriging believer of
 * optim_tools/MOBO.py:344-356 and the sequential updates of reliability_analysis/akmcs.py:127-139 and
 * optim_tools/SOBO.py:137-145 do when they re-run likelihood(KrigInfo['Theta'], KrigInfo, mode='all') at unchanged
 * hyper-parameters.  `next` is a model of n + 1 samples whose first n rows are the samples of `prev` (the caller
 * checks that), `prev` holds the fit of the same x (kb200_fit_state or kb200_fit_append).  The factor of the bordered
 * matrix is the old factor plus one row (O(n^2): one triangular solve with the old factor) instead of a new O(n^3)
 * factorisation; outputs, installed predictor and error reporting are those of kb200_fit_state on `next`.
 * Fails (< 0, nothing installed) if the hyper-parameters decoded from x differ from those of prev's fit or the two
 * models do not match; the caller then runs kb200_fit_state. */
int kb200_fit_append(kb200_model* next, kb200_model* prev, const double* x, int nbhyp, int trainvar,
                     double nugget_log10, double* U, double* Psi, double* BE, double* sigma2,
                     double* nll, int* info, double* theta_log10, double* nugget_out, double* wgkf);

/* Install a predictor from an existing KrigInfo (Theta, wgkf, U, BE, SigmaSqr), as
 * prediction() reads them (prediction.py:148-154).  U is n x n row-major upper. */
int kb200_predictor_load(kb200_model* m, const double* theta_log10, const double* wgkf,
                         const double* U, const double* BE, double sigma2);

/* prediction(x, KrigInfo, predtypes) for npred raw points (prediction.py:71-323):
 * standardise -> psi -> mean -> SSqr -> acquisition values.  Output pointers are npred
 * doubles each (NULL if not wanted, must match `want`).  stats[0] = #(SSqr == 0),
 * stats[1] = 1 if any EI term is non-zero (needed for the batch-level shortcuts of
 * prediction.py:273-274, 287-290).  ymin = min(y) (:272), limit/limit_ge = KrigInfo
 * ['limit'], ['limittype'] in ('>','>=') (:299-308). */
int kb200_predict(kb200_model* m, const double* x, long long npred, int normtype,
                  const double* norm_a, const double* norm_b, unsigned want, double ymin,
                  double limit, int limit_ge, double* pred, double* ssqr, double* s, double* ei,
                  double* poi, double* pof, double* ufun, unsigned long long* stats);
int kb200_predict_dev(kb200_model* m, const double* d_x, long long npred, int normtype,
                      const double* norm_a, const double* norm_b, unsigned want, double ymin,
                      double limit, int limit_ge, double* d_pred, double* d_ssqr, double* d_s,
                      double* d_ei, double* d_poi, double* d_pof, double* d_ufun,
                      unsigned long long* stats, void* stream);

/* AK-MCS learning / stopping quantities over a Monte-Carlo population without shipping the
 * npred predictions back (reliability_analysis/akmcs.py:267-297): U = abs(G)/s with its minimum and
 * first arg-min (lfucalc :281-285), counts[0] = #(G <= 0) (pfcalc :267-271), counts[1] =
 * #(G - 1.96 s <= 0), counts[2] = #(G + 1.96 s <= 0) (stopcrit :287-297); G = 'pred', s = 's'. */
int kb200_akmcs_reduce(kb200_model* m, const double* x, long long npred, int normtype,
                       const double* norm_a, const double* norm_b, double* min_u,
                       long long* argmin_u, unsigned long long* counts);

/* The same with the population already in HBM (d_x: [npred][d] device doubles): an AK-MCS run predicts the SAME
 * Monte-Carlo population after every added sample (akmcs.py:95-106, 144-155), so the caller uploads it once and every
 * iteration is this call -- no 8 d npred bytes over PCIe per iteration.  Runs on `stream` (NULL: the model's own),
 * returns the scalars on the host (blocking). */
int kb200_akmcs_reduce_dev(kb200_model* m, const double* d_x, long long npred, int normtype,
                           const double* norm_a, const double* norm_b, double* min_u,
                           long long* argmin_u, unsigned long long* counts, void* stream);

/* Saltelli sums for Sobol indices without shipping predictions back
 * (sensitivity_analysis/sobol_ind.py:45-141): A, B are the two [N][d] sample matrices (host), set q
 * builds C_q = B with the columns k where from_a[q*d + k] != 0 taken from A (one column: first /
 * total order :108-109; two columns: second order :152-154).  With ya, yb, yc_q the 'pred' outputs
 * on A, B, C_q:  base = {sum ya, sum ya^2, sum yb, sum yb^2},  ya_yc[q] = sum ya*yc_q,
 * yb_yc[q] = sum yb*yc_q.  The caller forms fo_2, denom and the indices (:84-85, :136-139, :180-182);
 * under data parallelism every rank passes its row shard and the sums are added. */
int kb200_sobol_sums(kb200_model* m, const double* A, const double* B, long long N, int normtype,
                     const double* norm_a, const double* norm_b, const unsigned char* from_a, int nsets,
                     double* base, double* ya_yc, double* yb_yc);

/* The same with A, B already in HBM ([N][d] device doubles each); runs on `stream` (NULL: the model's own), the sums
 * come back to the host (blocking). */
int kb200_sobol_sums_dev(kb200_model* m, const double* d_A, const double* d_B, long long N, int normtype,
                         const double* norm_a, const double* norm_b, const unsigned char* from_a, int nsets,
                         double* base, double* ya_yc, double* yb_yc, void* stream);
/* The same with A, B GENERATED on the device as the reference draws them (sensitivity_analysis/sobol_ind.py:24-29 via
 * misc/sampling/samplingplan.py:13-35): one unscrambled Sobol plan of 2 d columns, A = columns 0..d-1, B = columns
 * d..2d-1, rows i0 .. i0+N-1 of the sequence (the reference drops the all-zero row 0: i0 = 1; a rank that owns rows
 * [r0, r1) of the plan passes i0 = 1 + r0), scaled to [lo, hi] (2 d doubles each, realval :37-51; NULL: unit cube).
 * dirnum: [2 d][bits] direction numbers v_k,b of the host generator the caller would otherwise use (e.g. scipy's
 * qmc.Sobol(2 d, scramble=False)._sv), point i = xor of v_k,b over the set bits b of i ^ (i >> 1), times 2^-bits:
 * bit-identical to that generator, order-free.  Nothing but the sums crosses PCIe. */
int kb200_sobol_sums_gen(kb200_model* m, const unsigned long long* dirnum, int bits, long long i0, long long N,
                         const double* lo, const double* hi, int normtype, const double* norm_a, const double* norm_b,
                         const unsigned char* from_a, int nsets, double* base, double* ya_yc, double* yb_yc);
/* The generator alone: d_out [N][dims] device doubles = rows i0 .. i0+N-1 of the same sequence.  No model handle. */
int kb200_sobol_points_dev(int device, const unsigned long long* dirnum, int bits, int dims, long long i0, long long N,
                           const double* lo, const double* hi, double* d_out, void* stream);

/* 2-objective expected hyper-volume improvement of npop candidates over a Pareto front
 * (optim_tools/ehvi/exi2d.py:7-84 as called by EHVIcomputation.py:57 and :169-175): pareto is
 * [k][2] row-major (any order, k <= 1022), ref the reference point, candidate i has means
 * mu[2 i], mu[2 i + 1] and spreads s[2 i], s[2 i + 1] (the reference passes the kriging variance
 * SSqr here, EHVIcomputation.py:176; the library takes whatever it is given).  out[i] =
 * sign * exi2d(pareto, ref, mu_i, s_i); ehvicalc_vec uses sign = -1.  Host pointers.  No model
 * handle: the front is the only state. */
int kb200_ehvi2d(int device, const double* pareto, int k, const double* ref, const double* mu,
                 const double* s, long long npop, double sign, double* out);
/* Same with device pointers and element stride `inc` between candidates (inc = 2 with mu1 = mu0 + 1
 * for the [npop][2] layout, inc = 1 for one array per objective, e.g. the outputs of two
 * kb200_predict_dev calls); d_pareto is a device copy of the front; runs on `stream`. */
int kb200_ehvi2d_dev(int device, const double* d_pareto, int k, const double* ref,
                     const double* d_mu0, const double* d_mu1, const double* d_s0, const double* d_s1,
                     long long inc, long long npop, double sign, double* d_out, void* stream);

/* 3-objective expected hyper-volume improvement of npop candidates: replaces the reference's C++ extension
 * kmac.ehvi3d_sliceupdate_multi(y_par, ref_point, mean_vector, std_dev)
 * (kadal/extern/HDYE_3D_Update/kmac_wrapper.cpp:73-118 -> ehvi_multi.cc:113-341) as called by
 * optim_tools/ehvi/EHVIcomputation.py:235.  MAXIMISATION convention like KMAC (the caller negates, :235):
 * front is [k][3] row-major, ref the lower corner, candidate i has means mean[3 i + c] and spreads
 * sdev[3 i + c].  As in the wrapper (:24-34) the front is fed point by point and a new point drops the
 * earlier points it dominates or equals.  out[i] = sign * EHVI_i.
 * mode 0 reproduces the reference exactly, including ehvi_multi.cc:296-308 leaving the candidate loop with
 * `break` once one candidate's partial improvement in a cell is below 1e-15 (later candidates of the batch
 * lose that cell: results depend on the order of the batch, and the reference's own test_multi.py pins it);
 * mode 1 evaluates every candidate as if it were alone.  k <= 1000 (the reference's static tables).
 * Host pointers.  No model handle. */
#define KB200_EHVI3D_REFERENCE 0
#define KB200_EHVI3D_INDEPENDENT 1
int kb200_ehvi3d(int device, const double* front, int k, const double* ref, const double* mean,
                 const double* sdev, long long npop, int mode, double sign, double* out);
/* Same with device pointers: candidate i, objective c at d_mean[i * inc + c * cinc] (inc = 3, cinc = 1 for
 * [npop][3]; inc = 1, cinc = npop for one plane per objective); d_front must already be mutually
 * non-dominated (no filter); runs on `stream`. */
int kb200_ehvi3d_dev(int device, const double* d_front, int k, const double* ref, const double* d_mean,
                     const double* d_sdev, long long inc, long long cinc, long long npop, int mode, double sign,
                     double* d_out, void* stream);

/* Leave-one-out predictions of the installed fit: loo[i] = prediction at sample i of the model
 * refitted without it at the same hyper-parameters (beta re-estimated) -- what
 * supports/krigloocv2.py:7-34 computes with n refits -- in closed form from the factor. */
int kb200_loocv(kb200_model* m, double* loo /* n */);
/* The 'fast' variant, supports/krigloocv.py:4-38: loo[i] = y[i] - (B^-1[:n,:n] y)[i] / B^-1[i][i] with the bordered matrix
 * B = [[sigma2 * Psi, F], [F^T, 1]] -- as written in the reference: a ONE in the corner and KrigInfo['Psi'], i.e. WITHOUT the
 * nugget.  The installed fit must therefore be that of Psi itself (kb200_fit_state with nugget_log10 = -300 on a model of the
 * same samples); sigma2 = KrigInfo['SigmaSqr'] of the trained model.  Ordinary Kriging (p = 1).  Closed form over the factor. */
int kb200_loocv_fast(kb200_model* m, double sigma2, double* loo /* n */);

/* Introspection used by the tests / bench. */
int kb200_model_dims(const kb200_model* m, int* n, int* d, int* p, int* ntheta, int* nkernel);
/* kernel launches issued by this model since creation (bench.py "gpu_launches") */
unsigned long long kb200_launch_count(const kb200_model* m);
/* debug / parity: assemble Psi (+ nugget if add_eps) for one candidate into row-major n x n */
int kb200_debug_psi(kb200_model* m, const double* x, int nbhyp, int trainvar,
                    double nugget_log10, int add_eps, double* Psi);


/* debug / parity: raw tile storage (tile-major layout of DESIGN.md section 3) of candidate `cand` of the
 * last likelihood batch: tiles_per_mat * 16384 doubles */
int kb200_debug_tiles(kb200_model* m, int cand, double* out);

/* debug / parity: the exponential the PREDICTION kernels use for their (non-positive) arguments (which = 0: libm's
 * algorithm and coefficients without its denormal path and branch, arguments clamped at -708) or libm's exp itself
 * (which = 1), element-wise on n host doubles.  The likelihood's Psi always uses libm's exp. */
int kb200_debug_exp(int device, const double* x, long long n, int which, double* y);

/* Per-kernel-class device timing with CUDA events on the launching stream (bench.py's
 * roofline): enable, run, then read accumulated milliseconds and launch counts. */
#define KB200_PROF_ASSEMBLE 0   /* corr_tile_kernel, Psi assembly            */
#define KB200_PROF_DIAG 1       /* chol_diag_kernel                          */
#define KB200_PROF_PANEL 2      /* chol_panel_kernel (factorisation)         */
#define KB200_PROF_FINALIZE 3   /* lik_finalize_kernel                       */
#define KB200_PROF_PSI 4        /* corr_tile_kernel, psi(X, x*) + mean       */
#define KB200_PROF_PREDSOLVE 5  /* chol_panel_kernel (prediction variance)   */
#define KB200_PROF_EPILOGUE 6   /* predict_epilogue_kernel                   */
#define KB200_PROF_OTHER 7
#define KB200_PROF_NCLASS 8
int kb200_profile_enable(kb200_model* m, int on);
int kb200_profile_get(kb200_model* m, double* ms, unsigned long long* counts, int reset);

/* Number of candidate groups (1..4, default 2 or $KB200_GROUPS) a likelihood batch is split into;
 * every group runs its launch sequence on its own stream so that launch tails overlap.  With 1
 * all kernels are serialised on the caller's stream (what the per-kernel profile wants). */
int kb200_set_groups(kb200_model* m, int ngroups);

/* The factorisation has two schedules (left-looking for populations that fill the machine, right-looking for a few
 * matrices) that round differently (~1e-10 relative); which one runs is decided by the number of candidates of the
 * call.  A rank that evaluates a SHARD of a population passes the size of the whole population here, so that every
 * candidate gets the same bits -- and the population the same arg-min -- as in a single-GPU call on all of it.
 * 0 (default): decide from the call's own P. */
int kb200_set_schedule_population(kb200_model* m, long long population);

/* Empirical FP64 roof of the device: register-resident DMMA (mode 0) or DFMA (mode 1)
 * loops, best of 5, in TFLOP/s. */
int kb200_fp64_probe(int device, int mode, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* KB200_H */
