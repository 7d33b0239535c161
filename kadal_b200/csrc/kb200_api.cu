// kadal_b200 — C ABI (include/kb200.h).  Host-side orchestration only: every numeric
// step is a CUDA kernel from kb200_kernels.cu; there is no CPU compute path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <thread>
#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges show up in nsys / ncu --nvtx, no library to link
#include "../../include/kb200.h"
#include "kb200_kernels.h"

using namespace kb200;

static thread_local std::string g_err;

static int fail(const char* where, cudaError_t e) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return -(int)e - 1000;
}
static int fail_arg(const std::string& msg) {
  g_err = msg;
  return -1;
}

#define CK(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return fail(#call, e__); \
  } while (0)

// NVTX range around a host-side phase (assemble / factorise / predict ...): `ncu --nvtx --nvtx-include "kb200_factorize/"`
struct Nvtx {
  explicit Nvtx(const char* name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};

// Device allocations are recycled through a small process-wide pool: the sequential flows of KADAL (Kriging believer,
// AK-MCS, SOBO) build a new model of n + 1 samples for every update, and ~25 cudaMalloc / cudaFree pairs per model cost
// more than the refit itself.  A buffer goes back to the pool only after the device is idle (what cudaFree would
// have waited for as well); the pool is bounded and emptied when an allocation fails.
struct DevPool {
  struct Item { void* p; size_t cap; int dev; };
  std::mutex mu;
  std::vector<Item> items;
  size_t bytes = 0;
  static constexpr size_t MAX_BYTES = (size_t)8 << 30;
  static constexpr size_t MAX_ITEMS = 256;
  void* take(size_t want, int dev, size_t* cap) {
    std::lock_guard<std::mutex> g(mu);
    int best = -1;
    for (int i = 0; i < (int)items.size(); ++i)
      if (items[i].dev == dev && items[i].cap >= want && items[i].cap <= 2 * want + 4096 &&
          (best < 0 || items[i].cap < items[best].cap))
        best = i;
    if (best < 0) return nullptr;
    void* p = items[best].p;
    *cap = items[best].cap;
    bytes -= items[best].cap;
    items.erase(items.begin() + best);
    return p;
  }
  bool give(void* p, size_t cap, int dev) {   // caller guarantees the device is idle
    std::lock_guard<std::mutex> g(mu);
    if (items.size() >= MAX_ITEMS || bytes + cap > MAX_BYTES) return false;
    items.push_back({p, cap, dev});
    bytes += cap;
    return true;
  }
  void flush(int dev) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = 0; i < items.size();) {
      if (items[i].dev == dev) {
        cudaFree(items[i].p);
        bytes -= items[i].cap;
        items.erase(items.begin() + i);
      } else {
        ++i;
      }
    }
  }
};
static DevPool g_pool;
static const bool g_pool_on = !(getenv("KB200_NO_POOL") && atoi(getenv("KB200_NO_POOL")) != 0);

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void drop(bool idle) {
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_pool_on) {
      if (!idle) cudaDeviceSynchronize();
      if (!g_pool.give(p, cap, dev)) cudaFree(p);
    } else {
      cudaFree(p);
    }
    p = nullptr;
    cap = 0;
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    drop(false);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t want = bytes + bytes / 8;
    if (g_pool_on) {
      size_t got = 0;
      if (void* q = g_pool.take(bytes, dev, &got)) {
        p = q;
        cap = got;
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      g_pool.flush(dev);
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    else p = nullptr;
    return e;
  }
  void release(bool idle = false) { drop(idle); }
  double* d() const { return static_cast<double*>(p); }
};

struct PredSlot {
  cudaStream_t stream = nullptr;
  DevBuf x, chunk, fdot, udot, ssq, out[7];
  // variance chunks rotate over the call's stream and up to two side streams ($KB200_PRED_STREAMS)
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  DevBuf chunk2[2], ssq2[2];
};

struct kb200_model {
  int device = 0;
  int n = 0, d = 0, p = 0, h = 0, ntheta = 0, nkernel = 0, nq = 0, nt = 0, npad = 0;
  long tiles_per_mat = 0;
  double nsamp = 0;
  int flags = 0;
  int sm_count = 148;
  int kid[MAXK];
  int cstride = 0;
  std::vector<double> hy, hF;
  DevBuf X, R, pls, pls2;
  // KPLS: hyper-parameter independent projected-distance planes (kb200_kpls.cu), built on first use
  DevBuf kplanes_sq, kplanes_abs;
  bool kplanes_ready = false, kplanes_on = true;
  int half_mode = -1;       // half-SM factorisation kernels (kb200_half.cu): -1 auto, 0 off, 1 on
  cudaStream_t stream = nullptr;
  // likelihood workspace
  DevBuf Rw;
  bool rl_lookahead = true;
  int rl_mode = -1;          // right-looking factorisation for large matrices: -1 auto (nt > 16), 0 off, 1 on
  long sched_pop = 0;        // > 0: choose the schedule as for a population of this size (kb200_set_schedule_population)
  DevBuf tiles, W, Z, cand, xdev, nll, info, logdet, sigma2, beta, btb, resid, rhs2, sol2;
  // predictor state
  bool has_pred = false;
  DevBuf pL, pW, palpha, pw, pcand, pscale;
  double pbeta = 0, psigma2 = 0, pbtb = 0;
  // state of the last kb200_fit_state / kb200_fit_append that a one-sample append continues from:
  // Z = L^-1 [y F] ([nq][npad]) and sum ln L_ii
  bool has_fit = false;
  DevBuf pZ, plogdet, appdiag, appvec, trwork;
  PredSlot slot[2];
  // candidate groups of one likelihood batch run on their own streams so that the tail of one
  // group's launch (partially filled last wave, long "next diagonal" CTAs) overlaps the other's
  static constexpr int MAXG = 8;
  int ngroups = 2;
  cudaStream_t gstream[MAXG] = {nullptr};
  // half-SM path: per group a side stream for the panel tiles below row j + 1, so that the diagonal
  // tile j + 1 (latency-bound POTF2) runs under them (factorize())
  cudaStream_t gside[MAXG] = {nullptr};
  cudaEvent_t ev_d[MAXG] = {nullptr}, ev_p[MAXG] = {nullptr}, ev_asm[MAXG] = {nullptr};
  // $KB200_PRIO: 1 = the diagonal kernels of every group run on a high-priority stream of their own (their CTAs
  // are placed as soon as a slot frees instead of behind the other group's queued panel CTAs); 2 = the group
  // streams themselves get descending priorities (group 0 never waits for group 1)
  int prio = 0;
  cudaStream_t gprio[MAXG] = {nullptr};
  cudaEvent_t ev_q[MAXG] = {nullptr}, ev_r[MAXG] = {nullptr};
  bool stagger = false;     // measured: 8.35 vs 8.28 ms at two groups, worse with more ($KB200_STAGGER=1 to try)
  bool dual = false;        // measured: no gain over two candidate groups at P = 512 (8.32 ms both), kept as an option ($KB200_DUAL=1)
  int dual_env = -1;
  cudaEvent_t ev_fork = nullptr, ev_join[MAXG] = {nullptr};
  int group_min = 32;   // fewest candidates per group
  DevBuf norm_a, norm_b, stats, akpart;
  DevBuf sobB, sobC, sobY, sobAcc, sobPart, sobMask, sobDir;
  unsigned long long launches = 0;
  size_t ws_cap_bytes = 0;
  // One call in flight per model: every entry point that touches the shared scratch (cand, stats, norm vectors,
  // prediction slots, workspaces) first makes its stream wait for the previous call's last work (ev_last, recorded
  // on last_stream) -- calls issued on different streams are serialised by the library, never concurrent.
  cudaEvent_t ev_last = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_pending = false;
  // optional per-kernel-class event timing (kb200_profile_*)
  bool prof_on = false;
  struct ProfRec { cudaEvent_t a, b; int cls; };
  std::vector<ProfRec> prof;
  double prof_ms[KB200_PROF_NCLASS] = {0};
  unsigned long long prof_cnt[KB200_PROF_NCLASS] = {0};
};

// RAII: time one launch on stream `st` when profiling is on, and count it
struct Timed {
  kb200_model* m;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  Timed(kb200_model* m_, int cls, cudaStream_t st_) : m(m_), st(st_) {
    m->launches++;
    if (!m->prof_on) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { b = nullptr; return; }
    cudaEventRecord(a, st);
    m->prof.push_back({a, b, cls});
  }
  ~Timed() {
    if (b) cudaEventRecord(b, st);
  }
};

// stream ordering between successive calls on one model (see kb200_model::ev_last)
static cudaError_t call_enter(kb200_model* m, cudaStream_t st) {
  if (m->last_pending && m->last_stream != st) return cudaStreamWaitEvent(st, m->ev_last, 0);
  return cudaSuccess;
}
static cudaError_t call_leave(kb200_model* m, cudaStream_t st) {   // asynchronous return: work may still be running on st
  cudaError_t e = cudaEventRecord(m->ev_last, st);
  m->last_stream = st;
  m->last_pending = true;
  return e;
}
static void call_idle(kb200_model* m) { m->last_pending = false; }  // the call synchronised everything it launched

static CorrSpec make_spec(const kb200_model* m) {
  CorrSpec sp;
  sp.d = m->d;
  sp.ntheta = m->ntheta;
  sp.nkernel = m->nkernel;
  sp.kpls = m->h > 0 ? 1 : 0;
  for (int i = 0; i < MAXK; ++i) sp.kid[i] = m->kid[i];
  sp.pls = m->pls.d();
  sp.pls2 = m->pls2.d();
  return sp;
}

extern "C" {

int kb200_version(void) { return 100; }
const char* kb200_last_error(void) { return g_err.c_str(); }
int kb200_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}
void* kb200_host_alloc(unsigned long long bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void kb200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int kb200_model_create(kb200_model** out, int device, int n, int d, int p, int h,
                       const double* X, const double* y, const double* F, const double* pls,
                       const int* kernel_ids, int nkernel, double nsamp, int flags) {
  if (!out || !X || !y || !F || !kernel_ids) return fail_arg("kb200_model_create: null argument");
  if (n < 1 || d < 1 || p < 1) return fail_arg("kb200_model_create: n, d, p must be >= 1");
  if (p + 1 > MAXQ) return fail_arg("kb200_model_create: at most 7 trend columns (p) are supported");
  if (nkernel < 1 || nkernel > MAXK - 1) return fail_arg("kb200_model_create: 1..7 kernels supported");
  if (h < 0 || (h > 0 && !pls)) return fail_arg("kb200_model_create: plscoeff missing");
  for (int i = 0; i < nkernel; ++i) {
    if (kernel_ids[i] < 0 || kernel_ids[i] > 3) return fail_arg("Kernel option not available!");
    if (h > 0 && kernel_ids[i] >= KB200_KERNEL_MATERN32)
      return fail_arg("KPLS for matern32/matern52 is not yet implemented");
  }
  const int red_len = h > 0 ? h : d;
  if (red_len > 128) return fail_arg("kb200_model_create: more than 128 length scales not supported");
  if ((size_t)(2 * 128 * (d | 1) + 2) * 8 > 227 * 1024)
    return fail_arg("kb200_model_create: input dimension d > 110 not supported yet");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail_arg("kb200: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail_arg("kb200_model_create: bad device index");
  CK(cudaSetDevice(device));

  kb200_model* m = new kb200_model();
  m->device = device;
  m->n = n; m->d = d; m->p = p; m->h = h;
  m->ntheta = h > 0 ? h : d;
  m->nkernel = nkernel;
  m->nq = p + 1;
  m->nt = (n + TB - 1) / TB;
  m->npad = m->nt * TB;
  m->tiles_per_mat = (long)m->nt * (m->nt + 1) / 2;
  m->nsamp = nsamp;
  m->flags = flags;
  for (int i = 0; i < MAXK; ++i) m->kid[i] = i < nkernel ? kernel_ids[i] : 0;
  m->cstride = cand_stride(m->ntheta, nkernel);
  m->hy.assign(y, y + n);
  m->hF.assign(F, F + (size_t)n * p);
  cudaError_t e;
#define MCK(call)                      \
  do {                                 \
    e = (call);                        \
    if (e != cudaSuccess) {            \
      int rc = fail(#call, e);         \
      kb200_model_destroy(m);          \
      return rc;                       \
    }                                  \
  } while (0)
  MCK(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
  MCK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  MCK(cudaEventCreateWithFlags(&m->ev_last, cudaEventDisableTiming));
  for (int s = 0; s < 2; ++s) MCK(cudaStreamCreateWithFlags(&m->slot[s].stream, cudaStreamNonBlocking));
  {
    const char* eg = getenv("KB200_GROUPS");
    // default policy (measured, profiles/r1k_groups.md): two groups for the C2 shape; large
    // matrices with small populations (C5: nt = 32, P ~ 16) have few CTAs per tile column and
    // gain from four groups of >= 4 candidates.  KB200_GROUPS / kb200_set_groups() override.
    m->ngroups = m->nt >= 16 ? 4 : 2;
    m->group_min = std::max(2, std::min(32, 128 / m->nt));
    if (eg) m->ngroups = std::max(1, std::min((int)kb200_model::MAXG, atoi(eg)));
    const char* em = getenv("KB200_GROUP_MIN");
    if (em) m->group_min = std::max(1, atoi(em));
    {
      const char* ep = getenv("KB200_PRIO");
      if (ep) m->prio = atoi(ep);
    }
    int prio_least = 0, prio_greatest = 0;
    MCK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    for (int g = 0; g < m->ngroups; ++g) {
      if (m->prio == 2)
        MCK(cudaStreamCreateWithPriority(&m->gstream[g], cudaStreamNonBlocking, std::min(prio_least, prio_greatest + g)));
      else
        MCK(cudaStreamCreateWithFlags(&m->gstream[g], cudaStreamNonBlocking));
      MCK(cudaEventCreateWithFlags(&m->ev_join[g], cudaEventDisableTiming));
      MCK(cudaEventCreateWithFlags(&m->ev_asm[g], cudaEventDisableTiming));
    }
    MCK(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  }
  MCK(m->X.reserve((size_t)n * d * 8));
  MCK(cudaMemcpy(m->X.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice));
  // R = [y F]^T, zero padded: [nq][npad]
  {
    std::vector<double> R((size_t)m->nq * m->npad, 0.0);
    for (int i = 0; i < n; ++i) {
      R[i] = y[i];
      for (int k = 0; k < p; ++k) R[(size_t)(k + 1) * m->npad + i] = F[(size_t)i * p + k];
    }
    MCK(m->R.reserve(R.size() * 8));
    MCK(cudaMemcpy(m->R.p, R.data(), R.size() * 8, cudaMemcpyHostToDevice));
  }
  if (h > 0) {
    std::vector<double> p2((size_t)d * h);
    for (size_t i = 0; i < p2.size(); ++i) p2[i] = pls[i] * pls[i];
    MCK(m->pls.reserve(p2.size() * 8));
    MCK(m->pls2.reserve(p2.size() * 8));
    MCK(cudaMemcpy(m->pls.p, pls, p2.size() * 8, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(m->pls2.p, p2.data(), p2.size() * 8, cudaMemcpyHostToDevice));
  }
  {
    const char* es = getenv("KB200_STAGGER");
    if (es) m->stagger = atoi(es) != 0;
    const char* ed = getenv("KB200_DUAL");
    if (ed) { m->dual_env = atoi(ed) != 0 ? 1 : 0; m->dual = m->dual_env == 1; }
    const char* ela = getenv("KB200_RL_LOOKAHEAD");
    if (ela) m->rl_lookahead = atoi(ela) != 0;
    const char* erl = getenv("KB200_RL");
    if (erl) m->rl_mode = atoi(erl) != 0 ? 1 : 0;
    const char* eh = getenv("KB200_HALF");
    if (eh) m->half_mode = atoi(eh) != 0 ? 1 : 0;
    const char* ek = getenv("KB200_KPLS_PLANES");
    if (ek && atoi(ek) == 0) m->kplanes_on = false;
  }
  MCK(m->stats.reserve(16));
  {
    size_t fr = 0, tot = 0;
    MCK(cudaMemGetInfo(&fr, &tot));
    const char* envc = getenv("KB200_WORKSPACE_GB");
    size_t cap = envc ? (size_t)(atof(envc) * (1ull << 30)) : (size_t)(fr * 0.6);
    m->ws_cap_bytes = std::max<size_t>(cap, (size_t)256 << 20);
  }
#undef MCK
  *out = m;
  return 0;
}

void kb200_model_destroy(kb200_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  DevBuf* bufs[] = {&m->X, &m->R, &m->pls, &m->pls2, &m->kplanes_sq, &m->kplanes_abs, &m->tiles, &m->W, &m->Z, &m->cand, &m->xdev,
                    &m->nll, &m->info, &m->logdet, &m->sigma2, &m->beta, &m->btb, &m->resid,
                    &m->rhs2, &m->sol2, &m->pL, &m->pW, &m->palpha, &m->pw, &m->pcand, &m->pscale, &m->pZ, &m->plogdet, &m->appdiag, &m->appvec, &m->trwork,
                    &m->norm_a, &m->norm_b, &m->stats, &m->akpart,
                    &m->Rw, &m->sobB, &m->sobC, &m->sobY, &m->sobAcc, &m->sobPart, &m->sobMask, &m->sobDir};
  cudaDeviceSynchronize();   // nothing of this model is in flight when its buffers go back to the pool
  for (DevBuf* b : bufs) b->release(true);
  for (int s = 0; s < 2; ++s) {
    PredSlot& sl = m->slot[s];
    sl.x.release(true); sl.chunk.release(true); sl.fdot.release(true); sl.udot.release(true); sl.ssq.release(true);
    for (auto& o : sl.out) o.release(true);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    for (int k = 0; k < 2; ++k) {
      if (sl.side[k]) cudaStreamDestroy(sl.side[k]);
      if (sl.ev_join[k]) cudaEventDestroy(sl.ev_join[k]);
      sl.chunk2[k].release(true); sl.ssq2[k].release(true);
    }
    if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
  }
  for (int g = 0; g < kb200_model::MAXG; ++g) {
    if (m->gstream[g]) cudaStreamDestroy(m->gstream[g]);
    if (m->gside[g]) cudaStreamDestroy(m->gside[g]);
    if (m->gprio[g]) cudaStreamDestroy(m->gprio[g]);
    if (m->ev_q[g]) cudaEventDestroy(m->ev_q[g]);
    if (m->ev_r[g]) cudaEventDestroy(m->ev_r[g]);
    if (m->ev_d[g]) cudaEventDestroy(m->ev_d[g]);
    if (m->ev_p[g]) cudaEventDestroy(m->ev_p[g]);
    if (m->ev_asm[g]) cudaEventDestroy(m->ev_asm[g]);
    if (m->ev_join[g]) cudaEventDestroy(m->ev_join[g]);
  }
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_last) cudaEventDestroy(m->ev_last);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int kb200_model_dims(const kb200_model* m, int* n, int* d, int* p, int* ntheta, int* nkernel) {
  if (!m) return fail_arg("null model");
  if (n) *n = m->n;
  if (d) *d = m->d;
  if (p) *p = m->p;
  if (ntheta) *ntheta = m->ntheta;
  if (nkernel) *nkernel = m->nkernel;
  return 0;
}
unsigned long long kb200_launch_count(const kb200_model* m) { return m ? m->launches : 0; }

}  // extern "C"

// ------------------------------------------------------------------ internal drivers
static int layout_from_nbhyp(const kb200_model* m, int nbhyp, int trainvar, int* layout) {
  const bool iso = m->flags & KB200_FLAG_ISO;
  const int base = (iso ? 1 : m->ntheta) + (trainvar ? 1 : 0);
  const int extra = nbhyp - base;
  if (iso && (trainvar || extra != 0)) return fail_arg("iso_gaussian supports only a single theta");
  if (extra == 0) {
    if (m->nkernel != 1) return fail_arg("Nugget setting is not available");
    *layout = 0;
  } else if (extra == 1) {
    if (m->nkernel != 1) return fail_arg("Nugget setting is not available");
    *layout = 1;
  } else if (extra == m->nkernel) {
    *layout = 2;
  } else if (extra == m->nkernel + 1) {
    *layout = 3;
  } else {
    return fail_arg("Nugget setting is not available");
  }
  return 0;
}

static int decode(kb200_model* m, const double* d_x, long P, int nbhyp, int trainvar,
                  double nugget_log10, double* d_cand, cudaStream_t st) {
  int layout = 0;
  int rc = layout_from_nbhyp(m, nbhyp, trainvar, &layout);
  if (rc) return rc;
  DecodeArgs a;
  a.x = d_x; a.P = P; a.nbhyp = nbhyp; a.ntheta = m->ntheta; a.nkernel = m->nkernel;
  a.iso = (m->flags & KB200_FLAG_ISO) ? 1 : 0;
  a.theta_is_linear = 0;
  a.trainvar = trainvar;
  a.layout = layout;
  a.eps_fixed = pow(10.0, nugget_log10);
  a.wgkf_given = nullptr;
  a.cand = d_cand;
  a.cand_stride = m->cstride;
  {
    Timed t(m, KB200_PROF_OTHER, st);
    CK(launch_decode(a, st));
  }
  return 0;
}

static CholArgs chol_args(kb200_model* m, double* tiles, double* W, double* Z, double* logdet, int* info) {
  CholArgs a;
  memset(&a, 0, sizeof(a));
  a.L = tiles; a.W8 = W; a.tiles_per_mat = m->tiles_per_mat;
  a.nt = m->nt; a.npad = m->npad; a.n = m->n;
  a.nq = m->nq; a.R = m->R.d(); a.Z = Z; a.logdet = logdet; a.info = info;
  a.naive = (m->flags & KB200_FLAG_NAIVE) ? 1 : 0;
  a.pred_mode = 0;
  return a;
}

// KPLS (kernelfunc.py:41-47, 58-64): the distance planes do not depend on the candidate
static bool kpls_planes_usable(const kb200_model* m) {
  return m->h > 0 && m->h < 8 && m->kplanes_on && m->tiles_per_mat <= 65535 && !(m->flags & KB200_FLAG_NAIVE);
}

static int kpls_planes_build(kb200_model* m, cudaStream_t st) {
  if (m->kplanes_ready) return 0;
  bool need_sq = false, need_abs = false;
  for (int k = 0; k < m->nkernel; ++k) (m->kid[k] == K_GAUSS ? need_sq : need_abs) = true;
  const size_t bytes = (size_t)m->h * m->tiles_per_mat * TILE_ELEMS * 8;
  if (need_sq) {
    CK(m->kplanes_sq.reserve(bytes));
    Timed t(m, KB200_PROF_OTHER, st);
    CK(launch_kpls_planes(m->X.d(), m->n, m->d, m->h, m->pls2.d(), 0, m->kplanes_sq.d(), m->tiles_per_mat, st));
  }
  if (need_abs) {
    CK(m->kplanes_abs.reserve(bytes));
    Timed t(m, KB200_PROF_OTHER, st);
    CK(launch_kpls_planes(m->X.d(), m->n, m->d, m->h, m->pls.d(), 1, m->kplanes_abs.d(), m->tiles_per_mat, st));
  }
  // the planes are read by the candidate-group streams: make them visible device-wide first
  CK(cudaStreamSynchronize(st));
  m->kplanes_ready = true;
  return 0;
}

// assemble + factorise Pb candidates whose parameter rows start at d_cand
static int assemble(kb200_model* m, const double* d_cand, int Pb, double* tiles, int add_eps, cudaStream_t st) {
  Nvtx range("kb200_assemble");
  if (kpls_planes_usable(m)) {
    int rc = kpls_planes_build(m, st);
    if (rc) return rc;
    KplsAsmArgs k;
    memset(&k, 0, sizeof(k));
    k.n = m->n; k.h = m->h; k.nkernel = m->nkernel; k.add_eps = add_eps;
    for (int i = 0; i < MAXK; ++i) k.kid[i] = m->kid[i];
    k.tiles_per_mat = m->tiles_per_mat;
    k.planes_sq = m->kplanes_sq.d(); k.planes_abs = m->kplanes_abs.d();
    k.cand = d_cand; k.cand_stride = m->cstride; k.out = tiles;
    Timed t(m, KB200_PROF_ASSEMBLE, st);
    CK(launch_kpls_assemble(k, Pb, st));
    return 0;
  }
  CorrArgs c;
  memset(&c, 0, sizeof(c));
  c.spec = make_spec(m);
  c.mode = CORR_ASSEMBLE;
  c.n = m->n; c.nt = m->nt; c.X = m->X.d();
  c.cand = d_cand; c.cand_stride = m->cstride;
  c.out = tiles; c.tiles_per_mat = m->tiles_per_mat; c.add_eps = add_eps;
  {
    Timed t(m, KB200_PROF_ASSEMBLE, st);
    CK(launch_corr(c, (int)m->tiles_per_mat, Pb, st));
  }
  return 0;
}

static int factorize(kb200_model* m, int Pb, double* tiles, double* W, double* Z, double* logdet,
                     int* info, cudaStream_t st, int g = 0, long Ptotal = 1) {
  Nvtx range("kb200_factorize");
  CholArgs a = chol_args(m, tiles, W, Z, logdet, info);
  a.fuse_diag = a.naive ? 0 : 1;
  // Two schedules on the half-SM kernels (kb200_half.cu, two CTAs per SM), chosen by the amount of parallel
  // work of the CALL (candidates of the whole batch x tiles per dimension), not of the chunk or stream group:
  //  * left-looking (one diagonal + one panel launch per tile column, GEMM of depth 128 j per tile): least
  //    HBM traffic, best when the batch fills the machine (C2, P = 512: 8.3 ms vs 9.6 right-looking);
  //  * right-looking (every trailing tile updated per column, one tile of depth each): with few matrices the
  //    left-looking chain of one CTA per candidate and column (GEMM, solve, SYRK, POTF2) is the bound, here
  //    the chain is POTF2 + one solve and the trailing matrix is parallel work (n = 1000, P = 16: 0.86 vs
  //    1.14 ms; n = 2048, P = 8: 2.1 vs 3.9; C5 n = 4000, P = 16: 15.7 vs 18.9; tools/sweep_paths.py).
  // Cross-over measured at about 32 candidates / P x nt ~ 512 (no gain below 3 tiles per dimension).  The two schedules round differently (NLL agrees to ~1e-10
  // relative); within one schedule results are bit-identical for any batch composition, chunking or grouping.
  // $KB200_RL = 0 / 1 and $KB200_HALF = 0 (full-SM kernels of the first design) force a path.
  const bool use_rl = !a.naive && m->half_mode != 0 &&
                      (m->rl_mode == 1 ||
                       (m->rl_mode < 0 && m->nt >= 3 && (Ptotal <= 32 || (long)Ptotal * m->nt < 512)));
  const bool use_half = !a.naive && !use_rl && m->half_mode != 0;
  if (use_rl) {
    a.fuse_diag = 0;
    a.Rw = m->Rw.d() + (Z - m->Z.d());
    const int nt = m->nt;
    a.rl_mode = 1;
    CK(launch_rl_rhs_update(a, -1, Pb, st));
    m->launches++;
    // Look-ahead on two streams: after the solve of column k the main stream updates only the next column
    // (tiles (i, k+1) and the RHS tile k+1) and goes on with POTF2 / solve of column k+1, while the side
    // stream updates the rest of the trailing matrix.  U_a(k+1) touches tiles that U_b(k) also updates:
    // the main stream waits for U_b(k) before it (ev_p); U_b(k) needs the solve of column k (ev_d).
    if (!m->gside[g]) {
      CK(cudaStreamCreateWithFlags(&m->gside[g], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&m->ev_d[g], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&m->ev_p[g], cudaEventDisableTiming));
    }
    cudaStream_t side = m->gside[g];
    const bool look = m->rl_lookahead;
    for (int k = 0; k < nt; ++k) {
      {
        Timed t(m, KB200_PROF_DIAG, st);
        a.rl_mode = 1;
        a.i0 = 0;
        CK(launch_chol_diag_half(a, k, Pb, st));
      }
      const int M = nt - 1 - k;
      if (M <= 0) break;
      {
        Timed t(m, KB200_PROF_PANEL, st);
        a.rl_mode = 1;
        CK(launch_chol_panel_half(a, k, 2 * M, Pb, st));
      }
      if (!look) {
        {
          Timed t(m, KB200_PROF_PANEL, st);
          a.rl_mode = 2;
          CK(launch_chol_panel_half(a, k, M * (M + 1), Pb, st));
        }
        CK(launch_rl_rhs_update(a, k, Pb, st));
        m->launches++;
        continue;
      }
      // U_a(k) and the main-stream RHS update touch what U_b(k-1) also updates: wait for it here (this wait
      // is enqueued before ev_p is re-recorded below, so it refers to U_b(k-1))
      if (k > 0) CK(cudaStreamWaitEvent(st, m->ev_p[g], 0));
      CK(cudaEventRecord(m->ev_d[g], st));                       // column k solved
      if (M > 1) {                                               // U_b(k): pairs from row / column k + 2 on
        CK(cudaStreamWaitEvent(side, m->ev_d[g], 0));
        CholArgs b = a;
        b.rl_mode = 2;
        b.i0 = 1;
        {
          Timed t(m, KB200_PROF_PANEL, side);
          CK(launch_chol_panel_half(b, k, (M - 1) * M, Pb, side));
        }
        CK(launch_rl_rhs_update(b, k, Pb, side));                // RHS tiles k + 2 ..
        m->launches++;
      }
      {                                                          // U_a(k): the strip b = k + 1, RHS tile k + 1
        Timed t(m, KB200_PROF_PANEL, st);
        a.rl_mode = 3;
        a.i0 = 0;
        CK(launch_chol_panel_half(a, k, 2 * M, Pb, st));
      }
      CK(launch_rl_rhs_update(a, k, Pb, st, 1));
      m->launches++;
      CK(cudaEventRecord(m->ev_p[g], side));                     // (an empty record when M == 1)
    }
    if (look && nt > 1) CK(cudaStreamWaitEvent(st, m->ev_p[g], 0));   // join
    return 0;
  }
  if (use_half) {
    a.fuse_diag = 0;
    const int nt = m->nt;
    if (!m->dual || nt < 3) {
      const bool hp = m->prio == 1 && nt >= 2;
      if (hp && !m->gprio[g]) {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&m->gprio[g], cudaStreamNonBlocking, greatest));
        CK(cudaEventCreateWithFlags(&m->ev_q[g], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&m->ev_r[g], cudaEventDisableTiming));
      }
      for (int j = 0; j < nt; ++j) {
        if (hp) {
          CK(cudaEventRecord(m->ev_q[g], st));
          CK(cudaStreamWaitEvent(m->gprio[g], m->ev_q[g], 0));
          {
            Timed t(m, KB200_PROF_DIAG, m->gprio[g]);
            CK(launch_chol_diag_half(a, j, Pb, m->gprio[g]));
          }
          CK(cudaEventRecord(m->ev_r[g], m->gprio[g]));
          CK(cudaStreamWaitEvent(st, m->ev_r[g], 0));
        } else {
          Timed t(m, KB200_PROF_DIAG, st);
          CK(launch_chol_diag_half(a, j, Pb, st));
        }
        if (j + 1 < nt) {
          Timed t(m, KB200_PROF_PANEL, st);
          CK(launch_chol_panel_half(a, j, 2 * (nt - 1 - j), Pb, st));
        }
      }
      return 0;
    }
    // Two streams per group.  Column j: tile (j+1, j) alone on the main stream ("P_a"), then the diagonal
    // tile j+1 behind it; the tiles below ("P_b") on the side stream.  D(j+1) -- SYRK plus the strictly
    // sequential POTF2 -- needs only tile (j+1, j), so it runs co-resident with the tensor-bound P_b(j)
    // CTAs of its own group instead of after them.  Dependencies: P_b(j) needs D(j) (L_jj) [ev_d];
    // P_a(j+1) needs the tiles of row j+2 written by P_b(<= j) [ev_p].
    if (!m->gside[g]) {
      CK(cudaStreamCreateWithFlags(&m->gside[g], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&m->ev_d[g], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&m->ev_p[g], cudaEventDisableTiming));
    }
    cudaStream_t side = m->gside[g];
    {
      Timed t(m, KB200_PROF_DIAG, st);
      CK(launch_chol_diag_half(a, 0, Pb, st));
    }
    for (int j = 0; j + 1 < nt; ++j) {
      CK(cudaEventRecord(m->ev_d[g], st));               // D(j) done (and everything before it on main)
      if (j + 2 < nt) {
        CK(cudaStreamWaitEvent(side, m->ev_d[g], 0));
        CholArgs b = a;
        b.i0 = 1;
        {
          Timed t(m, KB200_PROF_PANEL, side);
          CK(launch_chol_panel_half(b, j, 2 * (nt - 2 - j), Pb, side));
        }
        CK(cudaEventRecord(m->ev_p[g], side));
      }
      {
        Timed t(m, KB200_PROF_PANEL, st);
        CK(launch_chol_panel_half(a, j, 2, Pb, st));     // tile (j+1, j); its row was completed by P_b(j-1),
      }                                                  // which main waited for at the end of the last round
      {
        Timed t(m, KB200_PROF_DIAG, st);
        CK(launch_chol_diag_half(a, j + 1, Pb, st));
      }
      if (j + 2 < nt) CK(cudaStreamWaitEvent(st, m->ev_p[g], 0));   // before P_a(j+1) and as the final join
    }
    return 0;
  }
  for (int j = 0; j < m->nt; ++j) {
    if (j == 0 || !a.fuse_diag) {
      Timed t(m, KB200_PROF_DIAG, st);
      CK(launch_chol_diag(a, j, Pb, st));
    }
    if (j + 1 < m->nt) {
      Timed t(m, KB200_PROF_PANEL, st);
      CK(launch_chol_panel(a, j, m->nt - 1 - j, Pb, st));
    }
  }
  return 0;
}

static size_t per_matrix_bytes(const kb200_model* m) {
  return (size_t)m->tiles_per_mat * TILE_ELEMS * 8 + (size_t)m->nt * 1024 * 8 + (size_t)m->nq * m->npad * 8 + 64;
}

static int reserve_lik(kb200_model* m, long P, int Pb) {
  CK(m->tiles.reserve((size_t)Pb * m->tiles_per_mat * TILE_ELEMS * 8));
  CK(m->W.reserve((size_t)Pb * m->nt * 1024 * 8));
  CK(m->Z.reserve((size_t)Pb * m->nq * m->npad * 8));
  CK(m->Rw.reserve((size_t)Pb * m->nq * m->npad * 8));
  CK(m->cand.reserve((size_t)P * m->cstride * 8));
  CK(m->logdet.reserve((size_t)Pb * 8));
  return 0;
}

extern "C" {

int kb200_lik_batch_dev(kb200_model* m, const double* d_x, int P, int nbhyp, int trainvar,
                        double nugget_log10, double* d_nll, int* d_info, void* stream) {
  if (!m || !d_x || !d_nll || !d_info) return fail_arg("kb200_lik_batch_dev: null argument");
  if (P < 1) return 0;
  Nvtx range("kb200_lik_batch");
  CK(cudaSetDevice(m->device));
  // stream == NULL: the model's own stream, and the call blocks until the results are complete (a caller that passes
  // no stream has nothing to order later work against); otherwise asynchronous on the caller's stream.
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  CK(call_enter(m, st));
  // Small models (the whole matrix fits in shared memory): one CTA per candidate assembles, factorises and solves in
  // shared memory -- nothing of Psi ever reaches HBM (lik_small_kernel; $KB200_NO_SMALL_LIK=1 forces the tile path).
  const bool no_small = getenv("KB200_NO_SMALL_LIK") && atoi(getenv("KB200_NO_SMALL_LIK")) != 0;   // (read per call: A/B tests)
  const int Nsmall = (no_small || (m->flags & KB200_FLAG_NAIVE)) ? 0 : lik_small_pad(m->n, m->d);
  if (Nsmall > 0) {
    const int Pb = std::min(P, 1 << 16);
    CK(m->Z.reserve((size_t)Pb * m->nq * m->npad * 8));
    CK(m->cand.reserve((size_t)P * m->cstride * 8));
    CK(m->logdet.reserve((size_t)Pb * 8));
    int rc = decode(m, d_x, P, nbhyp, trainvar, nugget_log10, m->cand.d(), st);
    if (rc) return rc;
    for (int p0 = 0; p0 < P; p0 += Pb) {
      const int pb = std::min(Pb, P - p0);
      CorrArgs c;
      memset(&c, 0, sizeof(c));
      c.spec = make_spec(m);
      c.mode = CORR_ASSEMBLE;
      c.n = m->n; c.nt = m->nt; c.X = m->X.d();
      c.cand = m->cand.d() + (size_t)p0 * m->cstride; c.cand_stride = m->cstride;
      c.add_eps = 1;
      CholArgs a = chol_args(m, nullptr, nullptr, m->Z.d(), m->logdet.d(), d_info + p0);
      {
        Nvtx r2("kb200_lik_small");
        Timed t(m, KB200_PROF_DIAG, st);
        CK(launch_lik_small(c, a, pb, st));
      }
      FinArgs f;
      memset(&f, 0, sizeof(f));
      f.Z = m->Z.d(); f.logdet = m->logdet.d(); f.info = d_info + p0;
      f.cand = c.cand; f.cand_stride = m->cstride;
      f.nq = m->nq; f.npad = m->npad; f.n = m->n;
      f.trainvar = trainvar; f.nsamp = m->nsamp;
      f.nll = d_nll + p0;
      {
        Timed t(m, KB200_PROF_FINALIZE, st);
        CK(launch_finalize(f, pb, st));
      }
    }
    if (!stream) {
      CK(cudaStreamSynchronize(st));
      call_idle(m);
    } else {
      CK(call_leave(m, st));
    }
    return 0;
  }
  int Pb = (int)std::min<size_t>((size_t)P, std::max<size_t>(1, m->ws_cap_bytes / per_matrix_bytes(m)));
  Pb = std::min(Pb, 32768);
  int rc = reserve_lik(m, P, Pb);
  if (rc) return rc;
  rc = decode(m, d_x, P, nbhyp, trainvar, nugget_log10, m->cand.d(), st);
  if (rc) return rc;
  CK(cudaMemsetAsync(d_info, 0, (size_t)P * sizeof(int), st));
  for (int p0 = 0; p0 < P; p0 += Pb) {
    const int pb = std::min(Pb, P - p0);
    const int G = std::max(1, std::min(m->ngroups, pb / m->group_min));
    if (G > 1) CK(cudaEventRecord(m->ev_fork, st));
    for (int g = 0; g < G; ++g) {
      const int o0 = (int)((long)pb * g / G), o1 = (int)((long)pb * (g + 1) / G), pg = o1 - o0;
      cudaStream_t sg = G > 1 ? m->gstream[g] : st;
      if (G > 1) CK(cudaStreamWaitEvent(sg, m->ev_fork, 0));
      const double* cand = m->cand.d() + (size_t)(p0 + o0) * m->cstride;
      double* tiles = m->tiles.d() + (size_t)o0 * m->tiles_per_mat * TILE_ELEMS;
      double* W8 = m->W.d() + (size_t)o0 * m->nt * 1024;
      double* Z = m->Z.d() + (size_t)o0 * m->nq * m->npad;
      double* logdet = m->logdet.d() + o0;
      // Optional stagger (off: no gain measured): group g assembles only after group g-1 has, so that the
      // groups stay about one tile column apart.
      if (G > 1 && g > 0 && m->stagger) CK(cudaStreamWaitEvent(sg, m->ev_asm[g - 1], 0));
      rc = assemble(m, cand, pg, tiles, 1, sg);
      if (rc) return rc;
      if (G > 1 && m->stagger) CK(cudaEventRecord(m->ev_asm[g], sg));
      rc = factorize(m, pg, tiles, W8, Z, logdet, d_info + p0 + o0, sg, g, m->sched_pop > 0 ? m->sched_pop : (long)P);
      if (rc) return rc;
      FinArgs f;
      memset(&f, 0, sizeof(f));
      f.Z = Z; f.logdet = logdet; f.info = d_info + p0 + o0;
      f.cand = cand; f.cand_stride = m->cstride;
      f.nq = m->nq; f.npad = m->npad; f.n = m->n;
      f.trainvar = trainvar; f.nsamp = m->nsamp;
      f.nll = d_nll + p0 + o0;
      {
        Timed t(m, KB200_PROF_FINALIZE, sg);
        CK(launch_finalize(f, pg, sg));
      }
      if (G > 1) {
        CK(cudaEventRecord(m->ev_join[g], sg));
        CK(cudaStreamWaitEvent(st, m->ev_join[g], 0));
      }
    }
  }
  if (!stream) {
    CK(cudaStreamSynchronize(st));
    call_idle(m);
  } else {
    CK(call_leave(m, st));
  }
  return 0;
}

int kb200_lik_batch(kb200_model* m, const double* x, int P, int nbhyp, int trainvar,
                    double nugget_log10, double* nll, int* info) {
  if (!m || !x || !nll || !info) return fail_arg("kb200_lik_batch: null argument");
  if (P < 1) return 0;
  CK(cudaSetDevice(m->device));
  CK(call_enter(m, m->stream));
  CK(m->xdev.reserve((size_t)P * nbhyp * 8));
  CK(m->nll.reserve((size_t)P * 8));
  CK(m->info.reserve((size_t)P * 4));
  CK(cudaMemcpyAsync(m->xdev.p, x, (size_t)P * nbhyp * 8, cudaMemcpyHostToDevice, m->stream));
  int rc = kb200_lik_batch_dev(m, m->xdev.d(), P, nbhyp, trainvar, nugget_log10, m->nll.d(),
                               static_cast<int*>(m->info.p), m->stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(nll, m->nll.p, (size_t)P * 8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(info, m->info.p, (size_t)P * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  call_idle(m);
  return 0;
}

int kb200_debug_psi(kb200_model* m, const double* x, int nbhyp, int trainvar,
                    double nugget_log10, int add_eps, double* Psi) {
  if (!m || !x || !Psi) return fail_arg("kb200_debug_psi: null argument");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  CK(call_enter(m, st));
  int rc = reserve_lik(m, 1, 1);
  if (rc) return rc;
  CK(m->xdev.reserve((size_t)nbhyp * 8));
  CK(cudaMemcpyAsync(m->xdev.p, x, (size_t)nbhyp * 8, cudaMemcpyHostToDevice, st));
  rc = decode(m, m->xdev.d(), 1, nbhyp, trainvar, nugget_log10, m->cand.d(), st);
  if (rc) return rc;
  rc = assemble(m, m->cand.d(), 1, m->tiles.d(), add_eps, st);
  if (rc) return rc;
  CK(m->sol2.reserve((size_t)m->n * m->n * 8));
  CK(launch_export(m->tiles.d(), m->n, m->sol2.d(), 1, st));
  m->launches++;
  CK(cudaMemcpyAsync(Psi, m->sol2.p, (size_t)m->n * m->n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// predictor constants from Z-space quantities: alpha = L^-T resid, w = L^-T b
static int install_predictor(kb200_model* m, const double* tiles, const double* W, const double* d_resid,
                             const double* d_b, const double* d_cand_row, double beta, double sigma2,
                             double btb, cudaStream_t st) {
  const size_t tb = (size_t)m->tiles_per_mat * TILE_ELEMS * 8, wb = (size_t)m->nt * 1024 * 8;
  CK(m->pL.reserve(tb));
  CK(m->pW.reserve(wb));
  CK(m->pcand.reserve((size_t)m->cstride * 8));
  CK(m->rhs2.reserve((size_t)2 * m->npad * 8));
  CK(m->sol2.reserve(std::max((size_t)2 * m->npad * 8, m->sol2.cap)));
  CK(m->palpha.reserve((size_t)m->npad * 8));
  CK(m->pw.reserve((size_t)m->npad * 8));
  if (tiles != m->pL.d()) CK(cudaMemcpyAsync(m->pL.p, tiles, tb, cudaMemcpyDeviceToDevice, st));
  if (W != m->pW.d()) CK(cudaMemcpyAsync(m->pW.p, W, wb, cudaMemcpyDeviceToDevice, st));
  if (d_cand_row != m->pcand.d())
    CK(cudaMemcpyAsync(m->pcand.p, d_cand_row, (size_t)m->cstride * 8, cudaMemcpyDeviceToDevice, st));
  if (m->nkernel == 1 && m->kid[0] == K_GAUSS) {
    CK(m->pscale.reserve((size_t)m->d * 8));
    CK(launch_pred_scale(m->pcand.d(), m->pls2.d(), m->d, m->ntheta, m->h > 0 ? 1 : 0, m->pscale.d(), st));
    m->launches++;
  }
  CK(cudaMemcpyAsync(m->rhs2.d(), d_resid, (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(m->rhs2.d() + m->npad, d_b, (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  BackArgs b;
  b.L = m->pL.d(); b.W8 = m->pW.d(); b.nt = m->nt; b.npad = m->npad; b.nq = 2;
  b.rhs = m->rhs2.d(); b.out = m->sol2.d();
  CK(m->trwork.reserve((size_t)MAXQ * m->npad * 8));
  b.work = m->trwork.d();
  CK(launch_backsolve(b, st));
  m->launches += trsv_launches(m->nt);
  CK(cudaMemcpyAsync(m->palpha.p, m->sol2.d(), (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(m->pw.p, m->sol2.d() + m->npad, (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  m->pbeta = beta; m->psigma2 = sigma2; m->pbtb = btb;
  m->has_pred = true;
  return 0;
}

// One-sample append: m's bordered matrix Psi + eps I is already assembled in m->tiles;
// prev holds the factor of its leading (n - 1) x (n - 1) block.  Leaves in m->tiles / W / Z / logdet / info what
// factorize() would leave, at O(n^2) instead of O(n^3) work.
static int append_factor(kb200_model* m, kb200_model* prev, cudaStream_t st) {
  const int np = prev->n;
  // appvec: [0, npad_prev) the new row of Psi, [npad_prev, 2 npad_prev) v = L^-1 psi
  CK(m->appvec.reserve((size_t)2 * prev->npad * 8));
  CK(m->appdiag.reserve(8));
  CK(m->trwork.reserve((size_t)MAXQ * m->npad * 8));
  CK(launch_append_extract(m->tiles.d(), np, prev->nt, m->appvec.d(), m->appdiag.d(), st));
  m->launches++;
  // the old factor is a prefix of the new tile storage (tile (ti, tj) at ti (ti + 1) / 2 + tj); the new row is
  // rewritten below, the rest of a new tile row keeps the assembly's identity padding
  CK(cudaMemcpyAsync(m->tiles.p, prev->pL.p, (size_t)prev->tiles_per_mat * TILE_ELEMS * 8, cudaMemcpyDeviceToDevice, st));
  BackArgs f;
  f.L = prev->pL.d(); f.W8 = prev->pW.d(); f.nt = prev->nt; f.npad = prev->npad; f.nq = 1;
  f.rhs = m->appvec.d(); f.out = m->appvec.d() + prev->npad; f.work = m->trwork.d();
  CK(launch_forwardsolve(f, st));
  m->launches += trsv_launches(prev->nt);
  CK(launch_append_row(m->tiles.d(), m->appvec.d() + prev->npad, np, prev->nt, m->nq, prev->npad, m->npad, prev->pZ.d(),
                       m->R.d(), m->Z.d(), prev->plogdet.d(), m->logdet.d(), static_cast<int*>(m->info.p),
                       m->appdiag.d(), st));
  m->launches++;
  CholArgs ca = chol_args(m, m->tiles.d(), m->W.d(), nullptr, nullptr, nullptr);
  CK(launch_w8_from_tiles(ca, 1, st));
  m->launches++;
  return 0;
}

static int fit_state_impl(kb200_model* m, kb200_model* prev, const double* x, int nbhyp, int trainvar,
                          double nugget_log10, double* U, double* Psi, double* BE, double* sigma2,
                          double* nll, int* info, double* theta_log10, double* nugget_out,
                          double* wgkf) {
  if (!m || !x) return fail_arg("kb200_fit_state: null argument");
  Nvtx range(prev ? "kb200_fit_append" : "kb200_fit_state");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  CK(call_enter(m, st));
  if (prev && prev != m) CK(call_enter(prev, st));   // the append reads prev's factor
  int rc = reserve_lik(m, 1, 1);
  if (rc) return rc;
  CK(m->xdev.reserve((size_t)nbhyp * 8));
  CK(m->nll.reserve(8));
  CK(m->info.reserve(4));
  CK(m->sigma2.reserve(8));
  CK(m->beta.reserve((size_t)m->p * 8));
  CK(m->btb.reserve(8));
  CK(m->resid.reserve((size_t)m->npad * 8));
  const size_t nn = (size_t)m->n * m->n * 8;
  CK(m->sol2.reserve(std::max(nn, (size_t)2 * m->npad * 8)));
  CK(cudaMemcpyAsync(m->xdev.p, x, (size_t)nbhyp * 8, cudaMemcpyHostToDevice, st));
  rc = decode(m, m->xdev.d(), 1, nbhyp, trainvar, nugget_log10, m->cand.d(), st);
  if (rc) return rc;
  CK(cudaMemsetAsync(m->info.p, 0, 4, st));
  // Psi without the nugget (KrigInfo['Psi'], likelihood_func.py:108), then + eps*I (:171)
  rc = assemble(m, m->cand.d(), 1, m->tiles.d(), 0, st);
  if (rc) return rc;
  if (Psi) {
    CK(launch_export(m->tiles.d(), m->n, m->sol2.d(), 1, st));
    m->launches++;
    CK(cudaMemcpyAsync(Psi, m->sol2.p, nn, cudaMemcpyDeviceToHost, st));
  }
  CK(launch_add_diag(m->tiles.d(), m->n, m->cand.d() + 4 * m->ntheta + m->nkernel, st));
  m->launches++;
  if (prev) {
    // the two fits must share every hyper-parameter: compare the decoded parameter rows bit for bit
    std::vector<double> a(m->cstride), b(m->cstride);
    CK(cudaMemcpyAsync(a.data(), m->cand.p, (size_t)m->cstride * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(b.data(), prev->pcand.p, (size_t)m->cstride * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (memcmp(a.data(), b.data(), (size_t)(m->cstride - 1) * 8) != 0)   // last entry: sigma2 of trainvar, not in Psi
      return fail_arg("kb200_fit_append: hyper-parameters differ from the fit installed in `prev`");
    rc = append_factor(m, prev, st);
  } else {
    rc = factorize(m, 1, m->tiles.d(), m->W.d(), m->Z.d(), m->logdet.d(), static_cast<int*>(m->info.p), st);
  }
  if (rc) return rc;
  FinArgs f;
  memset(&f, 0, sizeof(f));
  f.Z = m->Z.d(); f.logdet = m->logdet.d(); f.info = static_cast<int*>(m->info.p);
  f.cand = m->cand.d(); f.cand_stride = m->cstride;
  f.nq = m->nq; f.npad = m->npad; f.n = m->n; f.trainvar = trainvar; f.nsamp = m->nsamp;
  f.nll = m->nll.d(); f.sigma2 = m->sigma2.d(); f.beta = m->beta.d(); f.btb = m->btb.d();
  f.resid = m->resid.d();
  CK(launch_finalize(f, 1, st));
  m->launches++;
  if (U) {
    CK(launch_export(m->tiles.d(), m->n, m->sol2.d(), 0, st));
    m->launches++;
    CK(cudaMemcpyAsync(U, m->sol2.p, nn, cudaMemcpyDeviceToHost, st));
  }
  std::vector<double> hc(m->cstride), hbeta(m->p);
  double h_nll = 0, h_s2 = 0, h_btb = 0;
  int h_info = 0;
  CK(cudaMemcpyAsync(hc.data(), m->cand.p, (size_t)m->cstride * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hbeta.data(), m->beta.p, (size_t)m->p * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h_nll, m->nll.p, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h_s2, m->sigma2.p, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h_btb, m->btb.p, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h_info, m->info.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (BE) memcpy(BE, hbeta.data(), (size_t)m->p * 8);
  if (sigma2) *sigma2 = h_s2;
  if (nll) *nll = h_nll;
  if (info) *info = h_info;
  {
    const bool iso = m->flags & KB200_FLAG_ISO;
    int layout = 0;
    layout_from_nbhyp(m, nbhyp, trainvar, &layout);
    const int off = (iso ? 1 : m->ntheta) + (trainvar ? 1 : 0);
    if (theta_log10) for (int k = 0; k < m->ntheta; ++k) theta_log10[k] = iso ? x[0] : x[k];
    if (nugget_out) *nugget_out = (layout & 1) ? x[off] : nugget_log10;
    if (wgkf) for (int k = 0; k < m->nkernel; ++k) wgkf[k] = hc[4 * m->ntheta + k];
  }
  if (h_info == 0 && m->p == 1) {
    rc = install_predictor(m, m->tiles.d(), m->W.d(), m->resid.d(), m->Z.d() + m->npad,
                           m->cand.d(), hbeta[0], h_s2, h_btb, st);
    if (rc) return rc;
    // what a later one-sample append continues from
    CK(m->pZ.reserve((size_t)m->nq * m->npad * 8));
    CK(m->plogdet.reserve(8));
    CK(cudaMemcpyAsync(m->pZ.p, m->Z.p, (size_t)m->nq * m->npad * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(m->plogdet.p, m->logdet.p, 8, cudaMemcpyDeviceToDevice, st));
    m->has_fit = true;
    CK(cudaStreamSynchronize(st));
  } else {
    m->has_pred = false;
    m->has_fit = false;
  }
  return 0;
}

int kb200_fit_state(kb200_model* m, const double* x, int nbhyp, int trainvar,
                    double nugget_log10, double* U, double* Psi, double* BE, double* sigma2,
                    double* nll, int* info, double* theta_log10, double* nugget_out,
                    double* wgkf) {
  return fit_state_impl(m, nullptr, x, nbhyp, trainvar, nugget_log10, U, Psi, BE, sigma2, nll, info, theta_log10,
                        nugget_out, wgkf);
}

int kb200_fit_append(kb200_model* next, kb200_model* prev, const double* x, int nbhyp, int trainvar,
                     double nugget_log10, double* U, double* Psi, double* BE, double* sigma2,
                     double* nll, int* info, double* theta_log10, double* nugget_out, double* wgkf) {
  if (!next || !prev || !x) return fail_arg("kb200_fit_append: null argument");
  if (!prev->has_fit || !prev->has_pred)
    return fail_arg("kb200_fit_append: `prev` holds no fit (call kb200_fit_state or kb200_fit_append on it first)");
  if (next->n != prev->n + 1 || next->d != prev->d || next->p != prev->p || next->h != prev->h ||
      next->nkernel != prev->nkernel || next->device != prev->device ||
      ((next->flags ^ prev->flags) & KB200_FLAG_ISO) || memcmp(next->kid, prev->kid, sizeof(int) * next->nkernel) != 0)
    return fail_arg("kb200_fit_append: `next` must be `prev` with exactly one more sample (same d, trend, kernels, device)");
  if (next->p != 1) return fail_arg("kb200_fit_append: ordinary Kriging only (TrendOrder 0)");
  return fit_state_impl(next, prev, x, nbhyp, trainvar, nugget_log10, U, Psi, BE, sigma2, nll, info, theta_log10,
                        nugget_out, wgkf);
}

int kb200_predictor_load(kb200_model* m, const double* theta_log10, const double* wgkf,
                         const double* U, const double* BE, double sigma2) {
  if (!m || !theta_log10 || !wgkf || !U || !BE) return fail_arg("kb200_predictor_load: null argument");
  if (m->p != 1)
    return fail_arg("prediction with TrendOrder >= 1 is broken in the reference (prediction.py:250) and not supported");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  CK(call_enter(m, st));
  const size_t nn = (size_t)m->n * m->n * 8;
  CK(m->sol2.reserve(std::max(nn, (size_t)2 * m->npad * 8)));
  CK(m->pL.reserve((size_t)m->tiles_per_mat * TILE_ELEMS * 8));
  CK(m->pW.reserve((size_t)m->nt * 1024 * 8));
  CK(m->pcand.reserve((size_t)m->cstride * 8));
  CK(m->xdev.reserve((size_t)(m->ntheta + m->nkernel) * 8));
  // candidate row from Theta / wgkf (theta = 10**Theta, prediction.py:148)
  {
    std::vector<double> hx(m->ntheta + m->nkernel);
    for (int k = 0; k < m->ntheta; ++k) hx[k] = theta_log10[k];
    for (int k = 0; k < m->nkernel; ++k) hx[m->ntheta + k] = wgkf[k];
    CK(cudaMemcpyAsync(m->xdev.p, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.x = m->xdev.d(); a.P = 1; a.nbhyp = m->ntheta; a.ntheta = m->ntheta; a.nkernel = m->nkernel;
    a.iso = 0; a.trainvar = 0; a.layout = 0; a.eps_fixed = 0.0;
    a.wgkf_given = m->xdev.d() + m->ntheta;
    a.cand = m->pcand.d(); a.cand_stride = m->cstride;
    CK(launch_decode(a, st));
    m->launches++;
  }
  CK(cudaMemcpyAsync(m->sol2.p, U, nn, cudaMemcpyHostToDevice, st));
  CK(launch_import_U(m->sol2.d(), m->n, m->nt, m->pL.d(), st));
  m->launches++;
  CholArgs ca = chol_args(m, m->pL.d(), m->pW.d(), nullptr, nullptr, nullptr);
  CK(launch_w8_from_tiles(ca, 1, st));
  m->launches++;
  // forward solve of [y - F*BE | F]
  CK(m->rhs2.reserve((size_t)2 * m->npad * 8));
  std::vector<double> rhs((size_t)2 * m->npad, 0.0);
  for (int i = 0; i < m->n; ++i) {
    rhs[i] = m->hy[i] - m->hF[i] * BE[0];
    rhs[m->npad + i] = m->hF[i];
  }
  CK(cudaStreamSynchronize(st));  // sol2 (U staging) consumed before reuse below
  CK(cudaMemcpyAsync(m->rhs2.p, rhs.data(), rhs.size() * 8, cudaMemcpyHostToDevice, st));
  CK(m->Z.reserve((size_t)2 * m->npad * 8));
  BackArgs f;
  f.L = m->pL.d(); f.W8 = m->pW.d(); f.nt = m->nt; f.npad = m->npad; f.nq = 2;
  f.rhs = m->rhs2.d(); f.out = m->Z.d();
  CK(m->trwork.reserve((size_t)MAXQ * m->npad * 8));
  f.work = m->trwork.d();
  CK(launch_forwardsolve(f, st));
  m->launches += trsv_launches(m->nt);
  std::vector<double> hb(m->npad);
  CK(cudaMemcpyAsync(hb.data(), m->Z.d() + m->npad, (size_t)m->npad * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double btb = 0.0;
  for (int i = 0; i < m->n; ++i) btb = fma(hb[i], hb[i], btb);
  CK(m->resid.reserve((size_t)m->npad * 8));
  CK(cudaMemcpyAsync(m->resid.p, m->Z.d(), (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  CK(m->beta.reserve((size_t)m->npad * 8));
  CK(cudaMemcpyAsync(m->beta.p, m->Z.d() + m->npad, (size_t)m->npad * 8, cudaMemcpyDeviceToDevice, st));
  m->has_fit = false;
  int rc = install_predictor(m, m->pL.d(), m->pW.d(), m->resid.d(), m->beta.d(), m->pcand.d(), BE[0],
                             sigma2, btb, st);
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

// ------------------------------------------------------- bulk host -> device copies
// KADAL's callers hold their 10^7-point populations in ordinary (pageable) numpy arrays.  cudaMemcpyAsync from
// pageable memory is staged by the driver on the calling thread at ~10 GB/s (320 MB of C4 points: 31 ms against 13 ms
// of prediction kernels).  Large pageable sources are therefore copied by a few host threads into a small ring of
// pinned buffers and sent from there, so the copy of the next chunk proceeds at several times that rate while the
// kernels of the current one run.  Pinned sources and small copies go straight to cudaMemcpyAsync.
// $KB200_STAGE_THREADS (default 4, 0 = always the driver's path).
struct HostStager {
  static constexpr int RING = 3;
  static constexpr size_t SLOT = (size_t)32 << 20;
  void* buf[RING] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[RING] = {nullptr, nullptr, nullptr};
  bool used[RING] = {false, false, false};
  int next = 0, dev = -1, threads = 4;
  bool failed = false;
  std::mutex mu;
  bool init(int device) {
    if (failed) return false;
    if (buf[0]) return true;
    const char* e = getenv("KB200_STAGE_THREADS");
    if (e) threads = atoi(e);
    if (threads <= 0) { failed = true; return false; }
    threads = std::min(threads, 16);
    for (int i = 0; i < RING; ++i) {
      if (cudaHostAlloc(&buf[i], SLOT, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        for (int k = 0; k <= i; ++k) {
          if (buf[k]) cudaFreeHost(buf[k]);
          if (ev[k]) cudaEventDestroy(ev[k]);
          buf[k] = nullptr; ev[k] = nullptr;
        }
        failed = true;
        return false;
      }
    }
    dev = device;
    return true;
  }
};
// one ring per device, created on first use (the events and the pinned slots belong to that device's context)
static HostStager g_stagers[16];

static void parallel_memcpy(char* dst, const char* src, size_t bytes, int threads) {
  threads = (int)std::min<size_t>((size_t)threads, bytes >> 21);   // at least 2 MiB per thread
  if (threads <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t per = ((bytes / threads) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  size_t done_to = std::min(bytes, per);          // [0, per) is this thread's share
  for (int t = 1; t < threads; ++t) {
    const size_t lo = std::min(bytes, per * t), hi = (t == threads - 1) ? bytes : std::min(bytes, per * (t + 1));
    if (hi <= lo) continue;
    try {
      th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
    } catch (...) {                               // no more threads to be had: the calling thread copies the rest
      memcpy(dst + lo, src + lo, bytes - lo);
      break;
    }
  }
  memcpy(dst, src, done_to);
  for (auto& t : th) t.join();
}

static cudaError_t h2d_bulk(void* dst, const void* src, size_t bytes, int device, cudaStream_t st) {
  bool staged = bytes >= ((size_t)4 << 20);
  if (staged) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) != cudaSuccess) {
      cudaGetLastError();
      staged = false;
    } else if (at.type != cudaMemoryTypeUnregistered) {
      staged = false;                         // already pinned (or managed): the DMA engine reads it directly
    }
  }
  if (!staged || device < 0 || device >= 16) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  HostStager& g_stager = g_stagers[device];
  std::lock_guard<std::mutex> g(g_stager.mu);
  if (!g_stager.init(device)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  for (size_t off = 0; off < bytes; off += HostStager::SLOT) {
    const size_t len = std::min(HostStager::SLOT, bytes - off);
    const int k = g_stager.next;
    g_stager.next = (k + 1) % HostStager::RING;
    if (g_stager.used[k]) {
      cudaError_t e = cudaEventSynchronize(g_stager.ev[k]);
      if (e != cudaSuccess) return e;
    }
    parallel_memcpy(static_cast<char*>(g_stager.buf[k]), static_cast<const char*>(src) + off, len, g_stager.threads);
    cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + off, g_stager.buf[k], len, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(g_stager.ev[k], st);
    if (e != cudaSuccess) return e;
    g_stager.used[k] = true;
  }
  return cudaSuccess;
}

// --------------------------------------------------------------------- prediction
// variance streams of a bulk prediction: chunks of point tiles rotate over the call's stream and nstr - 1 side streams
static int pred_streams_for(const kb200_model* m) {
  static const int env = getenv("KB200_PRED_STREAMS") ? std::max(1, std::min(3, atoi(getenv("KB200_PRED_STREAMS")))) : 2;
  return (m->prof_on || m->half_mode == 0 || (m->flags & KB200_FLAG_NAIVE)) ? 1 : env;   // per-kernel timings want one stream
}

static long chunk_tiles_for(const kb200_model* m, long tiles_total = -1) {   // -1: the single-stream cap (host-path super-chunks, LOOCV)
  // Point tiles per variance chunk.  One wave of point tiles per launch left a quarter of the time in launch gaps and
  // tails (C4 pred+s: 17.8 ms at 1 wave, 15.4 at 8; with the half-SM panel kernel 14.2), and a chunk small enough to
  // stay in the 126 MB L2 is slower for the same reason (section 4 of DESIGN.md).  Late round 2: the chunks of a bulk
  // call rotate over TWO streams, so the tail of every launch of one chunk is filled by the other chunk's kernels (C3
  // 3.91e8 -> 4.2e8 pts/s, C4 pred + s 3.88e7 -> 4.1e7, C5 scoring 1.71e6 -> 1.82e6): the call is cut into an even number
  // of equal chunks of at most 16 waves / 2.5 GB of psi each ($KB200_CHUNK_WAVES, $KB200_PRED_STREAMS override).
  static const long env_waves = getenv("KB200_CHUNK_WAVES") ? std::max(1L, atol(getenv("KB200_CHUNK_WAVES"))) : 0;
  const long per_wave = 148L * std::max(1L, 4L / m->nt);
  const long bytes_per_wave = per_wave * m->nt * (long)TILE_ELEMS * 8;
  const int nstr = tiles_total < 0 ? 1 : pred_streams_for(m);
  const long wmax = nstr > 1 ? 16L : 8L;
  const long waves = env_waves ? env_waves : std::max(1L, std::min(wmax, (long)(2.5 * (1L << 30)) / bytes_per_wave));
  const long cap = per_wave * waves;
  if (nstr == 1 || tiles_total < 4 * 148L) return cap;                    // small batches: one chunk (split / right-looking paths)
  const long nchunks = (long)nstr * ((tiles_total + nstr * cap - 1) / (nstr * cap));
  return (tiles_total + nchunks - 1) / nchunks;
}

static int predict_core(kb200_model* m, PredSlot& sl, const double* d_x, long npts, int normtype,
                        unsigned want, double ymin, double limit, int limit_ge, double* outs[7],
                        cudaStream_t st) {
  Nvtx range("kb200_predict_core");
  const bool need_var = (want & ~KB200_OUT_PRED) != 0;
  CK(sl.fdot.reserve((size_t)npts * 8));
  CorrArgs c;
  memset(&c, 0, sizeof(c));
  c.spec = make_spec(m);
  c.mode = CORR_PREDICT;
  c.n = m->n; c.nt = m->nt; c.X = m->X.d();
  c.cand = m->pcand.d(); c.cand_stride = m->cstride;
  c.normtype = normtype; c.norm_a = m->norm_a.d(); c.norm_b = m->norm_b.d();
  c.pscale = (m->nkernel == 1 && m->kid[0] == K_GAUSS) ? m->pscale.d() : nullptr;
  c.alpha = m->palpha.d();
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.beta = m->pbeta; e.sigma2 = m->psigma2; e.btb = m->pbtb; e.ymin = ymin; e.limit = limit;
  e.limit_ge = limit_ge; e.need_var = need_var ? 1 : 0;
  e.stats = static_cast<unsigned long long*>(m->stats.p);
  if (!need_var) {
    c.xpts = d_x; c.npts = npts; c.out = nullptr; c.wvec = nullptr;
    c.fdot = sl.fdot.d(); c.udot = nullptr;
    const long tiles = (npts + TB - 1) / TB;
    if (m->nt >= 2 && tiles < m->sm_count) {
      // few points, many sample tiles: one CTA per (point tile, sample tile) pair, partial sums added by the epilogue
      const long pstride = (npts + 1) & ~1L;
      CK(sl.fdot.reserve((size_t)m->nt * pstride * 8));
      CorrArgs cc = c;
      cc.fdot = sl.fdot.d();
      cc.tj_split = 1;
      cc.part_stride = pstride;
      {
        Timed t(m, KB200_PROF_PSI, st);
        CK(launch_corr(cc, (int)tiles, m->nt, st));
      }
      e.npts = npts; e.fdot = sl.fdot.d(); e.pred = outs[0];
      e.nparts = m->nt; e.part_stride = pstride;
      {
        Timed t(m, KB200_PROF_EPILOGUE, st);
        CK(launch_epilogue(e, st));
      }
      return 0;
    }
    for (long t0 = 0; t0 < tiles; t0 += 1 << 20) {   // grid.x limit is 2^31-1; keep launches bounded
      const long tn = std::min<long>(1 << 20, tiles - t0);
      CorrArgs cc = c;
      cc.xpts = d_x + (size_t)t0 * TB * m->d;
      cc.npts = std::min<long>(npts - t0 * TB, tn * TB);
      cc.fdot = sl.fdot.d() + t0 * TB;
      {
        Timed t(m, KB200_PROF_PSI, st);
        CK(launch_corr(cc, (int)tn, 1, st));
      }
    }
    e.npts = npts; e.fdot = sl.fdot.d(); e.pred = outs[0];
    {
      Timed t(m, KB200_PROF_EPILOGUE, st);
      CK(launch_epilogue(e, st));
    }
    return 0;
  }
  // Small models (n <= 256, Gaussian, d <= 8) have a fused kernel with nothing of size n x npred in HBM (predict_small_kernel,
  // round 1).  Since the tiled path skips the identity padding of the last sample tile and solves sample tile 0 with the
  // resident-factor kernel (solve_rows_half_kernel) it is the faster one at every size measured (tools/time_pred_small.py,
  // pred + s over 4e6 points: n = 40 1.8e9 vs 1.0e9 pts/s, n = 128 9.1e8 vs 8.5e8, n = 200 (C3) 3.85e8 vs 3.60e8, n = 256
  // 2.97e8 vs 2.85e8), and it takes every kernel and any d: the fused kernel runs only on request ($KB200_FUSED_PREDICT=1).
  const char* ef = getenv("KB200_FUSED_PREDICT");                            // (read per call: A/B tests)
  const bool no_fused = !(ef && atoi(ef) != 0) || getenv("KB200_NO_FUSED_PREDICT") != nullptr;
  if (!no_fused && !(m->flags & KB200_FLAG_NAIVE) &&
      predict_small_supported(m->n, m->d, m->nkernel, m->kid[0], m->h > 0 ? 1 : 0)) {
    PredSmallArgs ps;
    memset(&ps, 0, sizeof(ps));
    ps.n = m->n; ps.d = m->d; ps.nt = m->nt; ps.ntheta = m->ntheta;
    ps.X = m->X.d(); ps.cand = m->pcand.d(); ps.pscale = m->pscale.d(); ps.L = m->pL.d(); ps.W8 = m->pW.d();
    ps.alpha = m->palpha.d(); ps.wvec = m->pw.d();
    ps.xpts = d_x; ps.npts = npts; ps.normtype = normtype;
    ps.norm_a = m->norm_a.d(); ps.norm_b = m->norm_b.d();
    ps.epi = e;
    ps.epi.npts = npts;
    ps.epi.pred = outs[0]; ps.epi.ssqr = outs[1]; ps.epi.s = outs[2]; ps.epi.ei = outs[3];
    ps.epi.poi = outs[4]; ps.epi.pof = outs[5]; ps.epi.ufun = outs[6];
    {
      Timed t(m, KB200_PROF_PREDSOLVE, st);
      CK(launch_predict_small(ps, m->sm_count, st));
    }
    return 0;
  }
  const long ct = std::min(chunk_tiles_for(m, (npts + TB - 1) / TB), (npts + TB - 1) / TB);
  const long cpts = ct * TB;
  CK(sl.chunk.reserve((size_t)ct * m->nt * TILE_ELEMS * 8));
  CK(sl.ssq.reserve((size_t)cpts * 8));
  CholArgs pa = chol_args(m, m->pL.d(), m->pW.d(), nullptr, nullptr, nullptr);
  pa.pred_mode = 1;
  pa.chunk = sl.chunk.d();
  pa.ssq = sl.ssq.d();
  // Few point tiles against many sample tiles (candidate scoring: EHVIcomputation.py:169-172 predicts n_pop ~ 10^2..10^3
  // candidates per generation on a model of n = 4000): the left-looking sweep has 2 x tiles CTAs per launch and a GEMM of
  // depth 128 j in each -- a handful of SMs busy for nt sequential launches.  Right-looking instead: solve tile j, then
  // update all nt - 1 - j later tiles of every point-tile row at once; and the psi kernel runs one CTA per (point tile,
  // sample tile) pair with per-sample-tile partial sums of psi^T alpha, psi^T w added in order by the epilogue.
  // The variance comes out bit-identical to the left-looking sweep (same products, accumulated from -psi in the same
  // order); the mean's psi^T alpha is summed per sample tile first (~1e-15 relative).  Chosen by batch size only.
  const int env_prl = getenv("KB200_PRED_RL") ? atoi(getenv("KB200_PRED_RL")) : -1;    // (read per call: A/B tests)
  const bool half_ok = m->half_mode != 0 && !pa.naive;
  // The chunks of a bulk call rotate over the call's stream and the side stream(s) (chunk_tiles_for): the launch tails of
  // one chunk's four kernels are filled by the other chunk's.  Fork / join by events: the call stays ordered on `st`.
  const int nstr = (npts > cpts && half_ok) ? pred_streams_for(m) : 1;
  const bool pingpong = nstr > 1;
  cudaStream_t st0 = st;
  if (pingpong) {
    if (!sl.ev_fork) CK(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
    CK(cudaEventRecord(sl.ev_fork, st0));
    for (int k = 0; k + 1 < nstr; ++k) {
      if (!sl.side[k]) {
        CK(cudaStreamCreateWithFlags(&sl.side[k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&sl.ev_join[k], cudaEventDisableTiming));
      }
      CK(sl.chunk2[k].reserve((size_t)ct * m->nt * TILE_ELEMS * 8));
      CK(sl.ssq2[k].reserve((size_t)cpts * 8));
      CK(cudaStreamWaitEvent(sl.side[k], sl.ev_fork, 0));
    }
    CK(sl.udot.reserve((size_t)npts * 8));
  }
  long ci = 0;
  for (long p0 = 0; p0 < npts; p0 += cpts, ++ci) {
    const long np = std::min(cpts, npts - p0);
    const int tiles = (int)((np + TB - 1) / TB);
    if (pingpong) {
      const int k = (int)(ci % nstr);
      st = k ? sl.side[k - 1] : st0;
      pa.chunk = k ? sl.chunk2[k - 1].d() : sl.chunk.d();
      pa.ssq = k ? sl.ssq2[k - 1].d() : sl.ssq.d();
    }
    const bool rl = !pingpong && half_ok && m->nt >= 3 && (env_prl >= 0 ? env_prl != 0 : 2 * tiles < 2 * m->sm_count);
    const bool split = !pingpong && (rl || (m->nt >= 2 && tiles < m->sm_count));
    const long pstride = (np + 1) & ~1L;
    CK(sl.udot.reserve((size_t)(split ? (size_t)m->nt * pstride : (size_t)npts) * 8));
    if (split) CK(sl.fdot.reserve(std::max((size_t)npts, (size_t)m->nt * pstride) * 8));
    CorrArgs cc = c;
    cc.xpts = d_x + (size_t)p0 * m->d; cc.npts = np; cc.out = pa.chunk;
    cc.wvec = m->pw.d();
    cc.fdot = split ? sl.fdot.d() : sl.fdot.d() + p0;
    cc.udot = split ? sl.udot.d() : sl.udot.d() + p0;
    cc.tj_split = split ? 1 : 0;
    cc.part_stride = pstride;
    cc.skip_pad = half_ok && half_kernels_skip_padding() ? 1 : 0;
    {
      Timed t(m, KB200_PROF_PSI, st);
      CK(launch_corr(cc, tiles, split ? m->nt : 1, st));
    }
    // right-looking sweep: solve tile 0, then ONE launch per tile column j: the CTAs of tile j + 1 apply column j's update and
    // go on to solve their (now final) tile, the CTAs of the later tiles only update ($KB200_PRED_RL_FUSED=0: two launches
    // per column, the first form; same bits either way)
    const bool rl_fused = rl && !(getenv("KB200_PRED_RL_FUSED") && atoi(getenv("KB200_PRED_RL_FUSED")) == 0);
    for (int j = 0; j < m->nt; ++j) {
      if (rl_fused) {
        Timed t(m, KB200_PROF_PREDSOLVE, st);
        if (j == 0) {
          pa.rl_mode = 1;
          CK(launch_chol_panel_half(pa, 0, 2 * tiles, 1, st));
        }
        if (j + 1 < m->nt) {
          pa.rl_mode = 4;
          CK(launch_chol_panel_half(pa, j, 2 * tiles * (m->nt - 1 - j), 1, st));
        }
        continue;
      }
      if (rl) {
        {
          Timed t(m, KB200_PROF_PREDSOLVE, st);
          pa.rl_mode = 1;
          CK(launch_chol_panel_half(pa, j, 2 * tiles, 1, st));
        }
        if (j + 1 < m->nt) {
          Timed t(m, KB200_PROF_PREDSOLVE, st);
          pa.rl_mode = 2;
          CK(launch_chol_panel_half(pa, j, 2 * tiles * (m->nt - 1 - j), 1, st));
        }
        continue;
      }
      Timed t(m, KB200_PROF_PREDSOLVE, st);
      if (half_ok) CK(launch_chol_panel_half(pa, j, 2 * tiles, 1, st));   // bit-identical to the full-SM kernel
      else CK(launch_chol_panel(pa, j, tiles, 1, st));
    }
    EpiArgs ee = e;
    ee.npts = np;
    ee.fdot = split ? sl.fdot.d() : sl.fdot.d() + p0;
    ee.udot = split ? sl.udot.d() : sl.udot.d() + p0;
    ee.nparts = split ? m->nt : 1;
    ee.part_stride = pstride;
    ee.ssq = pa.ssq;
    ee.pred = outs[0] ? outs[0] + p0 : nullptr;
    ee.ssqr = outs[1] ? outs[1] + p0 : nullptr;
    ee.s = outs[2] ? outs[2] + p0 : nullptr;
    ee.ei = outs[3] ? outs[3] + p0 : nullptr;
    ee.poi = outs[4] ? outs[4] + p0 : nullptr;
    ee.pof = outs[5] ? outs[5] + p0 : nullptr;
    ee.ufun = outs[6] ? outs[6] + p0 : nullptr;
    {
      Timed t(m, KB200_PROF_EPILOGUE, st);
      CK(launch_epilogue(ee, st));
    }
  }
  for (int k = 0; pingpong && k + 1 < nstr; ++k) {
    CK(cudaEventRecord(sl.ev_join[k], sl.side[k]));
    CK(cudaStreamWaitEvent(st0, sl.ev_join[k], 0));
  }
  return 0;
}

static int loocv_impl(kb200_model* m, double* loo, double fast_sigma2);
extern "C" int kb200_loocv(kb200_model* m, double* loo) { return loocv_impl(m, loo, 0.0); }
extern "C" int kb200_loocv_fast(kb200_model* m, double sigma2, double* loo) {
  if (!(sigma2 > 0.0)) return fail_arg("kb200_loocv_fast: sigma2 must be positive");
  return loocv_impl(m, loo, sigma2);
}
static int loocv_impl(kb200_model* m, double* loo, double fast_sigma2) {
  if (!m || !loo) return fail_arg("kb200_loocv: null argument");
  if (!m->has_pred) return fail_arg("kb200_loocv: no fit installed (call kb200_fit_state or kb200_predictor_load)");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  CK(call_enter(m, st));
  PredSlot& sl = m->slot[0];
  const long ct = std::min<long>(chunk_tiles_for(m), m->nt);
  const long cpts = ct * TB;
  CK(sl.chunk.reserve((size_t)ct * m->nt * TILE_ELEMS * 8));
  CK(sl.ssq.reserve((size_t)cpts * 8));
  CK(sl.out[0].reserve((size_t)m->npad * 8));
  CholArgs pa = chol_args(m, m->pL.d(), m->pW.d(), nullptr, nullptr, nullptr);
  pa.pred_mode = 1;
  pa.chunk = sl.chunk.d();
  pa.ssq = sl.ssq.d();
  for (long p0 = 0; p0 < m->n; p0 += cpts) {
    const long np = std::min<long>(cpts, m->n - p0);
    const int tiles = (int)((np + TB - 1) / TB);
    CK(launch_identity_chunk(sl.chunk.d(), tiles, m->nt, p0 / TB, st));
    m->launches++;
    for (int j = 0; j < m->nt; ++j) {
      Timed t(m, KB200_PROF_PREDSOLVE, st);
      if (m->half_mode != 0 && !pa.naive) CK(launch_chol_panel_half(pa, j, 2 * tiles, 1, st));   // bit-identical to the full-SM kernel
      else CK(launch_chol_panel(pa, j, tiles, 1, st));
    }
    CK(launch_loo(m->R.d(), m->palpha.d(), m->pw.d(), sl.ssq.d(), m->pbtb, m->n, p0, (int)np, sl.out[0].d(), st, fast_sigma2,
                  m->pbeta));
    m->launches++;
  }
  CK(cudaMemcpyAsync(loo, sl.out[0].p, (size_t)m->n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

static int predict_prepare(kb200_model* m, int normtype, const double* norm_a, const double* norm_b,
                           unsigned want, double* const outs[7], cudaStream_t st) {
  if (!m->has_pred) return fail_arg("kb200_predict: no predictor installed (call kb200_fit_state or kb200_predictor_load)");
  if (m->p != 1) return fail_arg("prediction with TrendOrder >= 1 is not supported (broken in the reference)");
  if (want == 0) return fail_arg("kb200_predict: nothing requested");
  for (int i = 0; i < 7; ++i)
    if (((want >> i) & 1u) && !outs[i]) return fail_arg("kb200_predict: requested output has a null pointer");
  if (normtype != KB200_NORM_NONE) {
    if (!norm_a || !norm_b) return fail_arg("kb200_predict: normalisation vectors missing");
    CK(m->norm_a.reserve((size_t)m->d * 8));
    CK(m->norm_b.reserve((size_t)m->d * 8));
    CK(cudaMemcpyAsync(m->norm_a.p, norm_a, (size_t)m->d * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->norm_b.p, norm_b, (size_t)m->d * 8, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(m->stats.p, 0, 16, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int kb200_predict_dev(kb200_model* m, const double* d_x, long long npred, int normtype,
                      const double* norm_a, const double* norm_b, unsigned want, double ymin,
                      double limit, int limit_ge, double* d_pred, double* d_ssqr, double* d_s,
                      double* d_ei, double* d_poi, double* d_pof, double* d_ufun,
                      unsigned long long* stats, void* stream) {
  if (!m || !d_x) return fail_arg("kb200_predict_dev: null argument");
  if (npred < 1) return 0;
  CK(cudaSetDevice(m->device));
  // stream == NULL: the model's own stream and a blocking call (see kb200_lik_batch_dev)
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  CK(call_enter(m, st));
  double* outs[7] = {d_pred, d_ssqr, d_s, d_ei, d_poi, d_pof, d_ufun};
  int rc = predict_prepare(m, normtype, norm_a, norm_b, want, outs, st);
  if (rc) return rc;
  for (int i = 0; i < 7; ++i)
    if (!((want >> i) & 1u)) outs[i] = nullptr;
  rc = predict_core(m, m->slot[0], d_x, (long)npred, normtype, want, ymin, limit, limit_ge, outs, st);
  if (rc) return rc;
  if (stats) CK(cudaMemcpyAsync(stats, m->stats.p, 16, cudaMemcpyDeviceToHost, st));
  if (stats || !stream) {
    CK(cudaStreamSynchronize(st));
    call_idle(m);
  } else {
    CK(call_leave(m, st));
  }
  return 0;
}

// x_is_device: the population is already resident in HBM (no copies, chunks run back to back on `user`, or on the
// model's stream when user is NULL); otherwise host memory, staged chunk by chunk on the two slot streams.
static int akmcs_core(kb200_model* m, const double* x, bool x_is_device, long long npred, int normtype,
                      const double* norm_a, const double* norm_b, double* min_u, long long* argmin_u,
                      unsigned long long* counts, cudaStream_t user) {
  if (!m || !x || !min_u || !argmin_u || !counts) return fail_arg("kb200_akmcs_reduce: null argument");
  if (npred < 1) return fail_arg("kb200_akmcs_reduce: empty population");
  Nvtx range("kb200_akmcs_reduce");
  CK(cudaSetDevice(m->device));
  const unsigned want = KB200_OUT_PRED | KB200_OUT_S;
  double* dummy[7] = {(double*)1, nullptr, (double*)1, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t base = x_is_device && user ? user : m->stream;
  CK(call_enter(m, base));       // predict_prepare synchronises `base`: the previous call is complete after it
  int rc = predict_prepare(m, normtype, norm_a, norm_b, want, dummy, base);
  if (rc) return rc;
  struct Part { double min_u; long long arg; unsigned long long c0, cm, cp; };
  static_assert(sizeof(Part) == AK_PARTIAL_BYTES, "AkPartial layout");
  const long ct = chunk_tiles_for(m) * TB;
  long sc = std::max<long>(ct, (1L << 19) / ct * ct);
  if (x_is_device) sc = std::max<long>(sc, (1L << 22) / ct * ct);   // nothing to overlap: fewer, larger launches
  const long nchunks = (npred + sc - 1) / sc;
  std::vector<Part> host((size_t)nchunks * AK_BLOCKS);
  CK(m->akpart.reserve((size_t)nchunks * AK_BLOCKS * sizeof(Part)));
  long it = 0;
  for (long p0 = 0; p0 < npred; p0 += sc, ++it) {
    PredSlot& sl = m->slot[x_is_device ? 0 : (it & 1)];
    cudaStream_t st = x_is_device ? base : sl.stream;
    const long np = std::min<long>(sc, npred - p0);
    const double* xs = x + (size_t)p0 * m->d;
    if (!x_is_device) {
      CK(sl.x.reserve((size_t)std::min<long>(sc, npred) * m->d * 8));
      CK(h2d_bulk(sl.x.p, xs, (size_t)np * m->d * 8, m->device, st));
      xs = sl.x.d();
    }
    double* outs[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    CK(sl.out[0].reserve((size_t)std::min<long>(sc, npred) * 8));
    CK(sl.out[2].reserve((size_t)std::min<long>(sc, npred) * 8));
    outs[0] = sl.out[0].d();
    outs[2] = sl.out[2].d();
    rc = predict_core(m, sl, xs, np, normtype, want, 0.0, 0.0, 1, outs, st);
    if (rc) return rc;
    CK(launch_akmcs_reduce(outs[0], outs[2], np, p0, static_cast<char*>(m->akpart.p) + (size_t)it * AK_BLOCKS * sizeof(Part),
                           AK_BLOCKS, st));
    m->launches++;
  }
  if (x_is_device) {
    CK(cudaMemcpyAsync(host.data(), m->akpart.p, host.size() * sizeof(Part), cudaMemcpyDeviceToHost, base));
    CK(cudaStreamSynchronize(base));
  } else {
    CK(cudaStreamSynchronize(m->slot[0].stream));
    CK(cudaStreamSynchronize(m->slot[1].stream));
    CK(cudaMemcpy(host.data(), m->akpart.p, host.size() * sizeof(Part), cudaMemcpyDeviceToHost));
  }
  call_idle(m);
  double bu = INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  unsigned long long c[3] = {0, 0, 0};
  for (const Part& p : host) {
    const bool un = p.min_u != p.min_u, bn = bu != bu;
    bool better;
    if (un != bn) better = un;
    else if (un) better = p.arg < bi;
    else better = p.min_u < bu || (p.min_u == bu && p.arg < bi);
    if (better) { bu = p.min_u; bi = p.arg; }
    c[0] += p.c0; c[1] += p.cm; c[2] += p.cp;
  }
  *min_u = bu;
  *argmin_u = bi;
  counts[0] = c[0]; counts[1] = c[1]; counts[2] = c[2];
  return 0;
}

int kb200_akmcs_reduce(kb200_model* m, const double* x, long long npred, int normtype,
                       const double* norm_a, const double* norm_b, double* min_u,
                       long long* argmin_u, unsigned long long* counts) {
  return akmcs_core(m, x, false, npred, normtype, norm_a, norm_b, min_u, argmin_u, counts, nullptr);
}

int kb200_akmcs_reduce_dev(kb200_model* m, const double* d_x, long long npred, int normtype,
                           const double* norm_a, const double* norm_b, double* min_u,
                           long long* argmin_u, unsigned long long* counts, void* stream) {
  return akmcs_core(m, d_x, true, npred, normtype, norm_a, norm_b, min_u, argmin_u, counts, (cudaStream_t)stream);
}

int kb200_predict(kb200_model* m, const double* x, long long npred, int normtype,
                  const double* norm_a, const double* norm_b, unsigned want, double ymin,
                  double limit, int limit_ge, double* pred, double* ssqr, double* s, double* ei,
                  double* poi, double* pof, double* ufun, unsigned long long* stats) {
  if (!m || !x) return fail_arg("kb200_predict: null argument");
  if (npred < 1) return 0;
  CK(cudaSetDevice(m->device));
  double* houts[7] = {pred, ssqr, s, ei, poi, pof, ufun};
  CK(call_enter(m, m->stream));       // predict_prepare synchronises m->stream: the previous call is complete after it
  int rc = predict_prepare(m, normtype, norm_a, norm_b, want, houts, m->stream);
  if (rc) return rc;
  // super-chunks alternate between two streams so that the H2D copy of chunk i+1 and the
  // D2H copy of chunk i-1 overlap the kernels of chunk i (pinned host memory required
  // for true overlap; pageable memory still works, staged by the driver).
  const long ct = chunk_tiles_for(m) * TB;
  long sc = std::max<long>(ct, (1L << 19) / ct * ct);
  // the first super-chunk is a single kernel chunk: the device starts after 1/3 .. 1/4 of the copy-in latency
  for (long p0 = 0, it = 0, np = 0; p0 < npred; p0 += np, ++it) {
    PredSlot& sl = m->slot[it & 1];
    cudaStream_t st = sl.stream;
    np = std::min<long>(it == 0 ? ct : sc, npred - p0);
    CK(sl.x.reserve((size_t)std::min<long>(sc, npred) * m->d * 8));
    CK(h2d_bulk(sl.x.p, x + (size_t)p0 * m->d, (size_t)np * m->d * 8, m->device, st));
    double* outs[7];
    for (int i = 0; i < 7; ++i) {
      outs[i] = nullptr;
      if ((want >> i) & 1u) {
        CK(sl.out[i].reserve((size_t)std::min<long>(sc, npred) * 8));
        outs[i] = sl.out[i].d();
      }
    }
    rc = predict_core(m, sl, sl.x.d(), np, normtype, want, ymin, limit, limit_ge, outs, st);
    if (rc) return rc;
    for (int i = 0; i < 7; ++i)
      if (outs[i]) CK(cudaMemcpyAsync(houts[i] + p0, outs[i], (size_t)np * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(m->slot[0].stream));
  CK(cudaStreamSynchronize(m->slot[1].stream));
  if (stats) CK(cudaMemcpy(stats, m->stats.p, 16, cudaMemcpyDeviceToHost));
  call_idle(m);
  return 0;
}

}  // extern "C"

extern "C" {

// ------------------------------------------------- Saltelli sums
// Where the two Saltelli sample matrices come from: host arrays (staged), device arrays (used in place), or the first
// 2 d columns of the unscrambled Sobol sequence generated on the device (sobol_ind.py:24-29: A = columns 0..d-1,
// B = columns d..2d-1 of one 2d-dimensional plan; samplingplan.py:18-19, 37-51: row 0 dropped, realval scaling).
struct SobolSource {
  int kind = 0;                 // 0 host, 1 device, 2 generated
  const double *A = nullptr, *B = nullptr;
  const unsigned long long* dirnum = nullptr;   // host [2 d][bits]
  int bits = 0;
  long long i0 = 0;             // sequence index of row 0 of this call
  const double *lo = nullptr, *hi = nullptr;    // host [2 d] or null (unit cube)
};

static int sobol_core(kb200_model* m, const SobolSource& src, long long N, int normtype,
                      const double* norm_a, const double* norm_b, const unsigned char* from_a, int nsets,
                      double* base, double* ya_yc, double* yb_yc, cudaStream_t user) {
  if (!m || !base || (nsets > 0 && (!from_a || !ya_yc || !yb_yc))) return fail_arg("kb200_sobol_sums: null argument");
  if (N < 1 || nsets < 0) return fail_arg("kb200_sobol_sums: empty sample");
  Nvtx range("kb200_sobol_sums");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = user ? user : m->stream;
  double* dummy[7] = {(double*)1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  CK(call_enter(m, st));
  int rc = predict_prepare(m, normtype, norm_a, norm_b, KB200_OUT_PRED, dummy, st);
  if (rc) return rc;
  static const int sc_log2 = getenv("KB200_SOBOL_CHUNK_LOG2") ? std::max(12, std::min(24, atoi(getenv("KB200_SOBOL_CHUNK_LOG2")))) : 18;
  const long sc = std::min<long>((long)N, 1L << sc_log2);
  const size_t cb = (size_t)sc * m->d * 8;
  PredSlot& sl = m->slot[0];
  if (src.kind != 1) {
    CK(sl.x.reserve(cb));
    CK(m->sobB.reserve(cb));
  }
  // The nsets + 2 prediction sets of a super-chunk are independent given A and B: the C_i sets alternate between the call's
  // stream and the second prediction slot's stream (own C buffer, y_C, partial sums and psi^T alpha scratch), so the launch
  // tails and the small compose / dot kernels of one set run under the psi kernel of the other ($KB200_SOBOL_STREAMS=1: off).
  static const bool two = !(getenv("KB200_SOBOL_STREAMS") && atoi(getenv("KB200_SOBOL_STREAMS")) < 2);
  const bool dual = two && nsets >= 2 && !m->prof_on;
  PredSlot& sl2 = m->slot[1];
  CK(m->sobC.reserve((dual ? 2 : 1) * cb));
  CK(m->sobY.reserve((size_t)4 * sc * 8));
  CK(m->sobAcc.reserve((size_t)(4 + 2 * nsets) * 8));
  CK(m->sobPart.reserve((size_t)4 * SOBOL_BLOCKS * 8));
  if (dual && !sl.ev_fork) CK(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
  if (dual && !sl.ev_join[0]) CK(cudaEventCreateWithFlags(&sl.ev_join[0], cudaEventDisableTiming));
  CK(m->sobMask.reserve((size_t)std::max(1, nsets) * m->d));
  if (nsets) CK(cudaMemcpyAsync(m->sobMask.p, from_a, (size_t)nsets * m->d, cudaMemcpyHostToDevice, st));
  if (src.kind == 2) {
    const int dims = 2 * m->d;
    CK(m->sobDir.reserve((size_t)dims * src.bits * 8 + (size_t)2 * dims * 8));
    CK(cudaMemcpyAsync(m->sobDir.p, src.dirnum, (size_t)dims * src.bits * 8, cudaMemcpyHostToDevice, st));
    std::vector<double> lohi((size_t)2 * dims);
    for (int k = 0; k < dims; ++k) { lohi[k] = src.lo ? src.lo[k] : 0.0; lohi[dims + k] = src.hi ? src.hi[k] : 1.0; }
    CK(cudaMemcpyAsync(static_cast<char*>(m->sobDir.p) + (size_t)dims * src.bits * 8, lohi.data(), lohi.size() * 8,
                       cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // lohi is a local
  }
  CK(cudaMemsetAsync(m->sobAcc.p, 0, (size_t)(4 + 2 * nsets) * 8, st));
  double *ya = m->sobY.d(), *yb = ya + sc, *yc = yb + sc, *acc = m->sobAcc.d(), *part = m->sobPart.d();
  const unsigned char* masks = static_cast<const unsigned char*>(m->sobMask.p);
  for (long p0 = 0; p0 < N; p0 += sc) {
    const long np = std::min<long>(sc, (long)N - p0);
    const double *dA, *dB;
    if (src.kind == 0) {
      CK(h2d_bulk(sl.x.p, src.A + (size_t)p0 * m->d, (size_t)np * m->d * 8, m->device, st));
      CK(h2d_bulk(m->sobB.p, src.B + (size_t)p0 * m->d, (size_t)np * m->d * 8, m->device, st));
      dA = sl.x.d(); dB = m->sobB.d();
    } else if (src.kind == 1) {
      dA = src.A + (size_t)p0 * m->d; dB = src.B + (size_t)p0 * m->d;
    } else {
      const int dims = 2 * m->d;
      const unsigned long long* dv = static_cast<const unsigned long long*>(m->sobDir.p);
      const double* lohi = reinterpret_cast<const double*>(dv + (size_t)dims * src.bits);
      const bool scaled = src.lo && src.hi;
      CK(launch_sobol_points(dv, src.bits, dims, src.i0 + p0, np, scaled ? lohi : nullptr, scaled ? lohi + dims : nullptr,
                             sl.x.d(), m->sobB.d(), m->d, st));
      m->launches++;
      dA = sl.x.d(); dB = m->sobB.d();
    }
    double* outs[7] = {ya, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    rc = predict_core(m, sl, dA, np, normtype, KB200_OUT_PRED, 0.0, 0.0, 1, outs, st);
    if (rc) return rc;
    outs[0] = yb;
    rc = predict_core(m, sl, dB, np, normtype, KB200_OUT_PRED, 0.0, 0.0, 1, outs, st);
    if (rc) return rc;
    CK(launch_sobol_dot(ya, nullptr, nullptr, np, 1, part, acc + 0, acc + 1, st));
    CK(launch_sobol_dot(yb, nullptr, nullptr, np, 1, part, acc + 2, acc + 3, st));
    m->launches += 4;
    if (dual) {
      CK(cudaEventRecord(sl.ev_fork, st));                     // A, B, y_A, y_B of this super-chunk are ready
      CK(cudaStreamWaitEvent(sl2.stream, sl.ev_fork, 0));
    }
    for (int q = 0; q < nsets; ++q) {
      const bool side = dual && (q & 1);
      cudaStream_t sq = side ? sl2.stream : st;
      double* Cq = m->sobC.d() + (side ? (size_t)sc * m->d : 0);
      double* ycq = side ? yc + sc : yc;
      double* partq = side ? part + 2 * SOBOL_BLOCKS : part;
      CK(launch_sobol_compose(dA, dB, masks + (size_t)q * m->d, np, m->d, Cq, sq));
      outs[0] = ycq;
      rc = predict_core(m, side ? sl2 : sl, Cq, np, normtype, KB200_OUT_PRED, 0.0, 0.0, 1, outs, sq);
      if (rc) return rc;
      CK(launch_sobol_dot(ya, yb, ycq, np, 0, partq, acc + 4 + 2 * q, acc + 5 + 2 * q, sq));
      m->launches += 3;
    }
    if (dual) {                                                // the next super-chunk overwrites what the side stream reads
      CK(cudaEventRecord(sl.ev_join[0], sl2.stream));
      CK(cudaStreamWaitEvent(st, sl.ev_join[0], 0));
    }
  }
  std::vector<double> h((size_t)4 + 2 * nsets);
  CK(cudaMemcpyAsync(h.data(), acc, h.size() * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  call_idle(m);
  for (int i = 0; i < 4; ++i) base[i] = h[i];
  for (int q = 0; q < nsets; ++q) { ya_yc[q] = h[4 + 2 * q]; yb_yc[q] = h[5 + 2 * q]; }
  return 0;
}

int kb200_sobol_sums(kb200_model* m, const double* A, const double* B, long long N, int normtype,
                     const double* norm_a, const double* norm_b, const unsigned char* from_a, int nsets,
                     double* base, double* ya_yc, double* yb_yc) {
  if (!A || !B) return fail_arg("kb200_sobol_sums: null argument");
  SobolSource src;
  src.kind = 0; src.A = A; src.B = B;
  return sobol_core(m, src, N, normtype, norm_a, norm_b, from_a, nsets, base, ya_yc, yb_yc, nullptr);
}

int kb200_sobol_sums_dev(kb200_model* m, const double* d_A, const double* d_B, long long N, int normtype,
                         const double* norm_a, const double* norm_b, const unsigned char* from_a, int nsets,
                         double* base, double* ya_yc, double* yb_yc, void* stream) {
  if (!d_A || !d_B) return fail_arg("kb200_sobol_sums_dev: null argument");
  SobolSource src;
  src.kind = 1; src.A = d_A; src.B = d_B;
  return sobol_core(m, src, N, normtype, norm_a, norm_b, from_a, nsets, base, ya_yc, yb_yc, (cudaStream_t)stream);
}

int kb200_sobol_sums_gen(kb200_model* m, const unsigned long long* dirnum, int bits, long long i0, long long N,
                         const double* lo, const double* hi, int normtype, const double* norm_a, const double* norm_b,
                         const unsigned char* from_a, int nsets, double* base, double* ya_yc, double* yb_yc) {
  if (!dirnum) return fail_arg("kb200_sobol_sums_gen: null argument");
  if (bits < 1 || bits > 64 || i0 < 0) return fail_arg("kb200_sobol_sums_gen: 1 <= bits <= 64, i0 >= 0");
  if (bits < 64 && (unsigned long long)(i0 + N) > (1ull << bits))
    return fail_arg("kb200_sobol_sums_gen: the sequence has only 2^bits points");
  SobolSource src;
  src.kind = 2; src.dirnum = dirnum; src.bits = bits; src.i0 = i0; src.lo = lo; src.hi = hi;
  return sobol_core(m, src, N, normtype, norm_a, norm_b, from_a, nsets, base, ya_yc, yb_yc, nullptr);
}

int kb200_sobol_points_dev(int device, const unsigned long long* dirnum, int bits, int dims, long long i0, long long N,
                           const double* lo, const double* hi, double* d_out, void* stream) {
  if (!dirnum || !d_out) return fail_arg("kb200_sobol_points_dev: null argument");
  if (bits < 1 || bits > 64 || dims < 1 || i0 < 0) return fail_arg("kb200_sobol_points_dev: bad bits / dims / i0");
  if (N < 1) return 0;
  if (bits < 64 && (unsigned long long)(i0 + N) > (1ull << bits))
    return fail_arg("kb200_sobol_points_dev: the sequence has only 2^bits points");
  CK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* dv = nullptr;
  CK(cudaMallocAsync((void**)&dv, (size_t)dims * bits * 8 + (size_t)2 * dims * 8, st));
  std::vector<double> lohi((size_t)2 * dims);
  for (int k = 0; k < dims; ++k) { lohi[k] = lo ? lo[k] : 0.0; lohi[dims + k] = hi ? hi[k] : 1.0; }
  double* dl = reinterpret_cast<double*>(dv + (size_t)dims * bits);
  cudaError_t e = cudaMemcpyAsync(dv, dirnum, (size_t)dims * bits * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dl, lohi.data(), lohi.size() * 8, cudaMemcpyHostToDevice, st);
  const bool scaled = lo && hi;
  if (e == cudaSuccess)
    e = launch_sobol_points(dv, bits, dims, i0, N, scaled ? dl : nullptr, scaled ? dl + dims : nullptr, d_out, nullptr, dims, st);
  cudaFreeAsync(dv, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // lohi / dirnum are the caller's and a local
  if (e != cudaSuccess) return fail("kb200_sobol_points_dev", e);
  return 0;
}

// ---------------------------------------------------------------- EHVI
static int ehvi_core(int device, const double* d_pareto, int k, const double* ref, const double* mu0,
                     const double* mu1, const double* s0, const double* s1, long inc, long npop, double sign,
                     double* d_out, cudaStream_t st) {
  if (k < 0 || k > 1022) return fail_arg("kb200_ehvi2d: 0 <= k <= 1022 Pareto points supported");
  const size_t k1 = (size_t)k + 1;
  double* ws = nullptr;     // S [2k] | c2 [k] | splus [(k+1)^2]
  CK(cudaMallocAsync((void**)&ws, (3 * (size_t)k + 2 + k1 * k1) * 8, st));
  double* S = ws;
  double* c2 = ws + 2 * (size_t)k + 1;
  double* splus = c2 + k + 1;
  int rc = 0;
  cudaError_t e = launch_ehvi_tables(d_pareto, k, ref[0], ref[1], S, c2, splus, st);
  if (e != cudaSuccess) rc = fail("ehvi tables", e);
  if (!rc) {
    EhviArgs a;
    a.k = k; a.r0 = ref[0]; a.r1 = ref[1]; a.S = S; a.c2 = c2; a.splus = splus;
    a.mu0 = mu0; a.mu1 = mu1; a.s0 = s0; a.s1 = s1; a.inc = inc; a.sign = sign; a.out = d_out;
    e = launch_ehvi_score(a, npop, st);
    if (e != cudaSuccess) rc = fail("ehvi score", e);
  }
  cudaFreeAsync(ws, st);
  return rc;
}

int kb200_ehvi2d_dev(int device, const double* d_pareto, int k, const double* ref,
                     const double* d_mu0, const double* d_mu1, const double* d_s0, const double* d_s1,
                     long long inc, long long npop, double sign, double* d_out, void* stream) {
  if (!ref || !d_mu0 || !d_mu1 || !d_s0 || !d_s1 || !d_out || (k > 0 && !d_pareto))
    return fail_arg("kb200_ehvi2d_dev: null argument");
  if (npop < 1) return 0;
  CK(cudaSetDevice(device));
  return ehvi_core(device, d_pareto, k, ref, d_mu0, d_mu1, d_s0, d_s1, (long)inc, (long)npop, sign, d_out,
                   (cudaStream_t)stream);
}

int kb200_ehvi2d(int device, const double* pareto, int k, const double* ref, const double* mu,
                 const double* s, long long npop, double sign, double* out) {
  if (!ref || !mu || !s || !out || (k > 0 && !pareto)) return fail_arg("kb200_ehvi2d: null argument");
  if (npop < 1) return 0;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail_arg("kb200: no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  double* buf = nullptr;    // pareto [2k] | mu [2 npop] | s [2 npop] | out [npop]
  const size_t np = (size_t)npop;
  cudaError_t e = cudaMalloc((void**)&buf, (2 * (size_t)k + 2 + 5 * np) * 8);
  if (e != cudaSuccess) { cudaStreamDestroy(st); return fail("cudaMalloc", e); }
  double* d_par = buf;
  double* d_mu = buf + 2 * (size_t)k + 2;
  double* d_s = d_mu + 2 * np;
  double* d_out = d_s + 2 * np;
  int rc = 0;
  auto CKL = [&](cudaError_t err, const char* what) { if (err != cudaSuccess && !rc) rc = fail(what, err); };
  if (k > 0) CKL(cudaMemcpyAsync(d_par, pareto, 2 * (size_t)k * 8, cudaMemcpyHostToDevice, st), "H2D pareto");
  CKL(cudaMemcpyAsync(d_mu, mu, 2 * np * 8, cudaMemcpyHostToDevice, st), "H2D mu");
  CKL(cudaMemcpyAsync(d_s, s, 2 * np * 8, cudaMemcpyHostToDevice, st), "H2D s");
  if (!rc) rc = ehvi_core(device, d_par, k, ref, d_mu, d_mu + 1, d_s, d_s + 1, 2, (long)npop, sign, d_out, st);
  if (!rc) CKL(cudaMemcpyAsync(out, d_out, np * 8, cudaMemcpyDeviceToHost, st), "D2H out");
  CKL(cudaStreamSynchronize(st), "sync");
  cudaFree(buf);
  cudaStreamDestroy(st);
  return rc;
}

// ---------------------------------------------------------------- 3-objective EHVI (KMAC replacement)
static int ehvi3_core(const double* d_front, int n, const double* ref, const double* d_mean, const double* d_sdev,
                      long inc, long cinc, long npop, int mode, double sign, double* d_out, cudaStream_t st) {
  if (n < 0 || n > 1000) return fail_arg("kb200_ehvi3d: 0 <= k <= 1000 front points supported (the reference's static tables)");
  if (mode != 0 && mode != 1) return fail_arg("kb200_ehvi3d: mode must be 0 (reference) or 1 (independent)");
  const size_t nn = (size_t)n * n, n1 = (size_t)n + 1;
  // Workspace: the front geometry (2 (n+1) n^2 doubles: 16 GB at the k = 1000 limit, 17 MB at k = 100) is built once;
  // the per-candidate terms ((9 + ysplit) (n+1) doubles per candidate) are bounded by a byte budget ($KB200_EHVI3_WS_MB,
  // default 4096) by scoring the candidates in chunks against the same tables -- in the independent mode, where
  // candidates do not interact.  The reference mode's early exit couples every candidate to the ones before it in the
  // batch (ehvi_multi.cc:296-308), so there the whole batch is one chunk, as in the reference.
  static const size_t budget = (size_t)(getenv("KB200_EHVI3_WS_MB") ? atof(getenv("KB200_EHVI3_WS_MB")) : 4096.0) << 20;
  long chunk = npop;
  if (mode == 1) {
    const size_t per_cand = (size_t)(9 + 16) * n1 * 8;
    chunk = (long)std::max<size_t>(1024, budget / per_cand);
    chunk = std::min(chunk, npop);
  }
  const size_t np = (size_t)chunk;
  // doubles: coord 3n | xlim nn | ylim nn | slice n1 nn | chunk n1 nn | E, G, EX 3 n1 np each | partial ysplit n1 np
  // then longs first 3 n1, ints rank 3n | hd nn
  const int ysplit = ehvi3_ysplit(n, chunk);
  const size_t nd = 3 * (size_t)n + 2 * nn + 2 * n1 * nn + (9 + (size_t)ysplit) * n1 * np;
  const size_t bytes = nd * 8 + 3 * n1 * 8 + (3 * (size_t)n + nn) * 4 + 64;
  unsigned char* ws = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&ws, bytes, st);
  if (e != cudaSuccess) return fail("kb200_ehvi3d workspace", e);
  Ehvi3Args a;
  a.n = n; a.npop = chunk; a.mode = mode; a.r0 = ref[0]; a.r1 = ref[1]; a.r2 = ref[2];
  a.front = d_front;
  double* p = reinterpret_cast<double*>(ws);
  a.coord = p; p += 3 * (size_t)n;
  a.xlim = p; p += nn;
  a.ylim = p; p += nn;
  a.slice = p; p += n1 * nn;
  a.chunk = p; p += n1 * nn;
  a.E = p; p += 3 * n1 * np;
  a.G = p; p += 3 * n1 * np;
  a.EX = p; p += 3 * n1 * np;
  a.partial = p; p += (size_t)ysplit * n1 * np;
  a.ysplit = ysplit;
  a.first = reinterpret_cast<long*>(p);
  a.rank = reinterpret_cast<int*>(a.first + 3 * n1);
  a.hd = a.rank + 3 * (size_t)n;
  a.inc = inc; a.cinc = cinc; a.sign = sign;
  int rc = 0;
  for (long c0 = 0; c0 < npop && !rc; c0 += chunk) {
    a.npop = std::min(chunk, npop - c0);
    a.mean = d_mean + c0 * inc; a.sdev = d_sdev + c0 * inc; a.out = d_out + c0;
    e = launch_ehvi3(a, st, c0 == 0 ? 3 : 2);
    if (e != cudaSuccess) rc = fail("ehvi3d", e);
  }
  cudaFreeAsync(ws, st);
  return rc;
}

int kb200_ehvi3d_dev(int device, const double* d_front, int k, const double* ref, const double* d_mean,
                     const double* d_sdev, long long inc, long long cinc, long long npop, int mode, double sign,
                     double* d_out, void* stream) {
  if (!ref || !d_mean || !d_sdev || !d_out || (k > 0 && !d_front)) return fail_arg("kb200_ehvi3d_dev: null argument");
  if (npop < 1) return 0;
  CK(cudaSetDevice(device));
  return ehvi3_core(d_front, k, ref, d_mean, d_sdev, (long)inc, (long)cinc, (long)npop, mode, sign, d_out,
                    (cudaStream_t)stream);
}

int kb200_ehvi3d(int device, const double* front, int k, const double* ref, const double* mean, const double* sdev,
                 long long npop, int mode, double sign, double* out) {
  if (!ref || !mean || !sdev || !out || (k > 0 && !front)) return fail_arg("kb200_ehvi3d: null argument");
  if (k < 0) return fail_arg("kb200_ehvi3d: k < 0");
  if (npop < 1) return 0;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail_arg("kb200: no CUDA device (there is no CPU fallback)");
  // the wrapper's marshalling (kmac_wrapper.cpp:24-34, 80-91): a new point drops the earlier points it dominates or equals
  std::vector<double> nd;
  nd.reserve(3 * (size_t)k);
  for (int i = 0; i < k; ++i) {
    const double* q = front + 3 * (size_t)i;
    size_t w = 0;
    for (size_t j = 0; j < nd.size(); j += 3) {
      const bool drop = q[0] >= nd[j] && q[1] >= nd[j + 1] && q[2] >= nd[j + 2];
      if (!drop) { nd[w] = nd[j]; nd[w + 1] = nd[j + 1]; nd[w + 2] = nd[j + 2]; w += 3; }
    }
    nd.resize(w);
    nd.insert(nd.end(), q, q + 3);
  }
  const int n = (int)(nd.size() / 3);
  CK(cudaSetDevice(device));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const size_t np = (size_t)npop;
  double* buf = nullptr;    // front [3n] | mean [3 npop] | sdev [3 npop] | out [npop]
  cudaError_t e = cudaMalloc((void**)&buf, (3 * (size_t)n + 3 + 7 * np) * 8);
  if (e != cudaSuccess) { cudaStreamDestroy(st); return fail("cudaMalloc", e); }
  double* d_front = buf;
  double* d_mean = buf + 3 * (size_t)n + 3;
  double* d_sdev = d_mean + 3 * np;
  double* d_out = d_sdev + 3 * np;
  int rc = 0;
  auto CKL = [&](cudaError_t err, const char* what) { if (err != cudaSuccess && !rc) rc = fail(what, err); };
  if (n > 0) CKL(cudaMemcpyAsync(d_front, nd.data(), 3 * (size_t)n * 8, cudaMemcpyHostToDevice, st), "H2D front");
  CKL(cudaMemcpyAsync(d_mean, mean, 3 * np * 8, cudaMemcpyHostToDevice, st), "H2D mean");
  CKL(cudaMemcpyAsync(d_sdev, sdev, 3 * np * 8, cudaMemcpyHostToDevice, st), "H2D sdev");
  if (!rc) rc = ehvi3_core(d_front, n, ref, d_mean, d_sdev, 3, 1, (long)npop, mode, sign, d_out, st);
  if (!rc) CKL(cudaMemcpyAsync(out, d_out, np * 8, cudaMemcpyDeviceToHost, st), "D2H out");
  CKL(cudaStreamSynchronize(st), "sync");
  cudaFree(buf);
  cudaStreamDestroy(st);
  return rc;
}

// debug: raw tile storage of candidate `cand` of the last likelihood batch (tiles_per_mat * 16384 doubles)
int kb200_debug_tiles(kb200_model* m, int cand, double* out) {
  if (!m || !out) return fail_arg("kb200_debug_tiles: null argument");
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, m->tiles.d() + (size_t)cand * m->tiles_per_mat * TILE_ELEMS,
                (size_t)m->tiles_per_mat * TILE_ELEMS * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int kb200_debug_exp(int device, const double* x, long long n, int which, double* y) {
  if (!x || !y) return fail_arg("kb200_debug_exp: null argument");
  if (n < 1) return 0;
  CK(cudaSetDevice(device));
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, (size_t)2 * n * 8));
  cudaError_t e = cudaMemcpy(d, x, (size_t)n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_debug_exp(d, (long)n, which, d + n, 0);
  if (e == cudaSuccess) e = cudaMemcpy(y, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail("kb200_debug_exp", e);
  return 0;
}

int kb200_set_groups(kb200_model* m, int ngroups) {
  if (!m) return fail_arg("null model");
  if (ngroups < 1 || ngroups > kb200_model::MAXG) return fail_arg("kb200_set_groups: 1..8 groups");
  CK(cudaSetDevice(m->device));
  for (int g = 0; g < ngroups; ++g) {
    if (!m->gstream[g]) CK(cudaStreamCreateWithFlags(&m->gstream[g], cudaStreamNonBlocking));
    if (!m->ev_join[g]) CK(cudaEventCreateWithFlags(&m->ev_join[g], cudaEventDisableTiming));
    if (!m->ev_asm[g]) CK(cudaEventCreateWithFlags(&m->ev_asm[g], cudaEventDisableTiming));
  }
  m->ngroups = ngroups;
  m->dual = m->dual_env == 1 && ngroups > 1;   // 1 group = everything serialised on one stream
  return 0;
}

int kb200_set_schedule_population(kb200_model* m, long long population) {
  if (!m) return fail_arg("null model");
  if (population < 0) return fail_arg("kb200_set_schedule_population: population >= 0 (0 = decide per call)");
  m->sched_pop = (long)population;
  return 0;
}

int kb200_profile_enable(kb200_model* m, int on) {
  if (!m) return fail_arg("null model");
  m->prof_on = on != 0;
  return 0;
}

static int prof_collect(kb200_model* m) {
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  for (auto& r : m->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      m->prof_ms[r.cls] += ms;
      m->prof_cnt[r.cls]++;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  m->prof.clear();
  return 0;
}

int kb200_profile_get(kb200_model* m, double* ms, unsigned long long* counts, int reset) {
  if (!m) return fail_arg("null model");
  int rc = prof_collect(m);
  if (rc) return rc;
  for (int i = 0; i < KB200_PROF_NCLASS; ++i) {
    if (ms) ms[i] = m->prof_ms[i];
    if (counts) counts[i] = m->prof_cnt[i];
    if (reset) { m->prof_ms[i] = 0; m->prof_cnt[i] = 0; }
  }
  return 0;
}

int kb200_fp64_probe(int device, int mode, double* tflops) {
  if (!tflops) return fail_arg("null argument");
  CK(cudaSetDevice(device));
  double* sink = nullptr;
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int blocks = 148 * 4, iters = 4096;
  CK(launch_fp64_probe(blocks, 64, mode, sink, 0));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a, 0));
    CK(launch_fp64_probe(blocks, iters, mode, sink, 0));
    CK(cudaEventRecord(b, 0));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    // per thread and iteration: DMMA 16 x (8*8*4*2)/32 flops, DFMA 32 x 2 flops
    const double per_thread = mode == 0 ? 16.0 * 512.0 / 32.0 : 64.0;
    const double fl = per_thread * iters * 512.0 * blocks;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  return 0;
}

}  // extern "C"
