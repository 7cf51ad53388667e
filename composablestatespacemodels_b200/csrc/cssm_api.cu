// cssm_api.cu -- host side of libcssm_gpu.so: the C ABI of include/cssm.h over the kernels of
// cssm_kernels.cuh.  One CUDA stream per filter handle, no global mutable state, no CPU fallback.
#include <cmath>
#include <cstdlib>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>

#include "cssm_kernels.cuh"
#include "cssm_series.cuh"
#include "cssm_forecast.cuh"

using namespace cssm;

namespace {

thread_local std::string g_err = "";

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? CSSM_ERR_NOMEM : CSSM_ERR_CUDA,                \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

struct HostLeaf {
  int sde_kind, dim, f_kind, period, harmonics;
  std::vector<double> m0, c0, phi, mu, sigma;
};
struct HostModel {
  std::vector<HostLeaf> leaves;
  int obs_kind, has_scale, step_mode, lgcp_precision, d, obs_df;
  double scale;
};

int copy_model(const cssm_model_desc_t* m, HostModel& out) {
  if (!m || m->n_leaves <= 0 || !m->leaves) return fail(CSSM_ERR_INVALID, "model descriptor: no leaves");
  if (m->obs_kind < CSSM_OBS_POISSON || m->obs_kind > CSSM_OBS_BETA) return fail(CSSM_ERR_INVALID, "model descriptor: unknown obs_kind");
  if (m->obs_kind == CSSM_OBS_STUDENT_T && !m->has_scale) return fail(CSSM_ERR_INVALID, "No scale parameter provided to Student T Model");
  if (m->obs_kind == CSSM_OBS_STUDENT_T && m->obs_df <= 0) return fail(CSSM_ERR_INVALID, "model descriptor: Student T model needs obs_df > 0");
  if (m->obs_kind == CSSM_OBS_ZIP && !m->has_scale)
    return fail(CSSM_ERR_INVALID, "Must provide probability parameter for zero inflated Poisson Model");
  if (m->step_mode != CSSM_STEP_EXACT && m->step_mode != CSSM_STEP_EULER) return fail(CSSM_ERR_INVALID, "model descriptor: unknown step_mode");
  if ((m->obs_kind == CSSM_OBS_NEGBIN || m->obs_kind == CSSM_OBS_NORMAL) && !m->has_scale)
    return fail(CSSM_ERR_INVALID, m->obs_kind == CSSM_OBS_NEGBIN ? "No scale parameter provided to Negativebinomial Model"
                                                                 : "Must provide SD parameter for LinearModel");
  out.leaves.clear();
  out.d = 0;
  for (int l = 0; l < m->n_leaves; ++l) {
    const cssm_leaf_t& L = m->leaves[l];
    HostLeaf h;
    h.sde_kind = L.sde_kind; h.dim = L.dim; h.f_kind = L.f_kind; h.period = L.period; h.harmonics = L.harmonics;
    if (L.dim <= 0) return fail(CSSM_ERR_INVALID, "model descriptor: leaf dimension must be positive");
    if (L.sde_kind < CSSM_SDE_BROWNIAN || L.sde_kind > CSSM_SDE_OU) return fail(CSSM_ERR_INVALID, "model descriptor: unknown sde_kind");
    if (L.f_kind == CSSM_F_SEASONAL && (L.dim != 2 * L.harmonics || L.period <= 0))
      return fail(CSSM_ERR_INVALID, "model descriptor: seasonal leaf needs dim == 2*harmonics and period > 0");
    if (!L.m0 || !L.c0 || !L.sigma) return fail(CSSM_ERR_INVALID, "model descriptor: m0/c0/sigma missing");
    if (L.sde_kind != CSSM_SDE_BROWNIAN && !L.mu) return fail(CSSM_ERR_INVALID, "model descriptor: mu missing");
    if (L.sde_kind == CSSM_SDE_OU && !L.phi) return fail(CSSM_ERR_INVALID, "model descriptor: phi missing");
    h.m0.assign(L.m0, L.m0 + L.dim);
    h.c0.assign(L.c0, L.c0 + L.dim);
    h.sigma.assign(L.sigma, L.sigma + L.dim);
    if (L.mu) h.mu.assign(L.mu, L.mu + L.dim);
    if (L.phi) h.phi.assign(L.phi, L.phi + L.dim);
    out.d += L.dim;
    out.leaves.push_back(h);
  }
  if (out.d > MAXD) return fail(CSSM_ERR_UNSUPPORTED, "total latent dimension exceeds 32");
  out.obs_kind = m->obs_kind; out.has_scale = m->has_scale; out.scale = m->scale;
  out.step_mode = m->step_mode; out.lgcp_precision = m->lgcp_precision;
  out.obs_df = m->obs_df;
  return CSSM_OK;
}

// f-coefficients at time t: gamma = sum_k C[k] x[k]   (model/Model.scala:184,217-225)
void f_coeffs(const HostModel& m, double t, double* C) {
  int k = 0;
  for (const HostLeaf& L : m.leaves) {
    for (int c = 0; c < L.dim; ++c) C[k + c] = 0.0;
    if (L.f_kind == CSSM_F_SEASONAL) {
      double frequency = 2 * M_PI / L.period;
      for (int a = 1; a <= L.harmonics; ++a) {
        C[k + 2 * (a - 1)] = std::cos(frequency * a * t);
        C[k + 2 * (a - 1) + 1] = std::sin(frequency * a * t);
      }
    } else {
      C[k] = 1.0;
    }
    k += L.dim;
  }
}

// per-step constants in fp64 (see StepArgs in cssm_kernels.cuh for the transition forms)
struct StepHost {
  double A[MAXD], M[MAXD], D[MAXD], S[MAXD], C[MAXD];
  double y, k0, k1, k2, k3;
  int has_obs;
};
void transition_consts(const HostModel& m, double dt, StepHost& s) {
  int k = 0;
  for (const HostLeaf& L : m.leaves)
    for (int c = 0; c < L.dim; ++c, ++k) {
      double A = 1.0, M = 0.0, D = 0.0, S;
      if (m.step_mode == CSSM_STEP_EXACT) {
        switch (L.sde_kind) {
          case CSSM_SDE_BROWNIAN: S = std::sqrt(L.sigma[c] * dt); break;                       // model/Sde.scala:114-123
          case CSSM_SDE_GEN_BROWNIAN: D = L.mu[c] * dt; S = std::sqrt(L.sigma[c] * dt); break;  // :86-95
          default: {                                                                            // :139-150
            double phi = L.phi[c], sigma = L.sigma[c];
            A = std::exp(-phi * dt);
            M = L.mu[c];
            S = std::sqrt((sigma * sigma / (phi * 2.0)) * (1.0 - std::exp(phi * -2.0 * dt)));
          }
        }
      } else {  // Euler-Maruyama, model/Sde.scala:30-43 with :82-84,:110-112,:158-162
        S = L.sigma[c] * std::sqrt(dt);
        switch (L.sde_kind) {
          case CSSM_SDE_BROWNIAN: D = 1.0 * dt; break;
          case CSSM_SDE_GEN_BROWNIAN: D = L.mu[c] * dt; break;
          default: A = 1.0 - L.phi[c] * dt; M = L.mu[c];
        }
      }
      s.A[k] = A; s.M[k] = M; s.D[k] = D; s.S[k] = S;
    }
}
void obs_consts(const HostModel& m, int has_obs, double y, StepHost& s) {
  s.y = y; s.has_obs = has_obs; s.k0 = s.k1 = s.k2 = s.k3 = 0.0;
  if (!has_obs) return;
  switch (m.obs_kind) {
    case CSSM_OBS_POISSON: { int k = (int)y; s.k0 = k; s.k1 = std::lgamma(k + 1.0); break; }
    case CSSM_OBS_NEGBIN: {
      int k = (int)y;
      double size = std::exp(m.scale);
      s.k0 = k; s.k1 = size;
      s.k2 = std::lgamma(size + k) - std::lgamma(k + 1.0) - std::lgamma(size);
      s.k3 = std::log(size);
      break;
    }
    case CSSM_OBS_NORMAL: { double v = std::exp(m.scale); s.k0 = v; s.k1 = std::log(std::sqrt(2 * M_PI)) + std::log(v); break; }
    case CSSM_OBS_BERNOULLI: s.k0 = (y == 1.0) ? 1.0 : 0.0; break;
    case CSSM_OBS_STUDENT_T: {  // 1/v * StudentsT(df).logPdf((y - eta)/v), model/Model.scala:154-160
      const double df = (double)m.obs_df;
      s.k0 = std::exp(m.scale);                                                                   // v
      s.k1 = df;
      s.k2 = std::lgamma((df + 1.0) / 2.0) - std::lgamma(df / 2.0) - 0.5 * std::log(M_PI * df);   // -logNormalizer
      s.k3 = (df + 1.0) / 2.0;
      break;
    }
    case CSSM_OBS_ZIP: {  // model/Model.scala:298-306
      const int k = (int)y;
      const double ev = std::exp(m.scale);
      s.k0 = k; s.k1 = std::lgamma(k + 1.0); s.k2 = ev / (1.0 + ev); s.k3 = std::log(1.0 + ev);
      break;
    }
    case CSSM_OBS_BETA:  // Beta(exp(-gamma), 1.0).logPdf(y) = (a-1) log y + (1-1) log(1-y) + log a, model/Model.scala:349-352
      s.k0 = std::log(y);
      s.k1 = 0.0 * std::log(1.0 - y);  // NaN for y == 1, as (b-1)*log(1-x) is in the reference
      break;
    default: break;
  }
}
template <typename real>
void to_args(const HostModel& m, const StepHost& h, StepArgs<real>& a) {
  std::memset(&a, 0, sizeof(a));
  for (int k = 0; k < m.d; ++k) {
    // x' = A*(x - M) + M + D + S*z  ==  A*x + B + S*z  with  B = M - A*M + D  (formed in fp64)
    a.A[k] = (real)h.A[k]; a.M[k] = (real)h.M[k]; a.D[k] = (real)(h.M[k] - h.A[k] * h.M[k] + h.D[k]); a.S[k] = (real)h.S[k]; a.C[k] = (real)h.C[k];
  }
  a.y = (real)h.y; a.k0 = (real)h.k0; a.k1 = (real)h.k1; a.k2 = (real)h.k2; a.k3 = (real)h.k3;
  a.d = m.d; a.obs_kind = m.obs_kind; a.has_obs = h.has_obs;
}

uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct SeriesStep {
  double t, dt;
  StepHost h;       // transition for dt (for LGCP: for one sub-step delta), C at time t
  long long n_sub;  // LGCP sub-steps (1 otherwise)
  size_t ctab_off;  // offset into the LGCP coefficient table (elements)
};

}  // namespace

struct cssm_filter {
  int device = 0, dtype = CSSM_F32, resample_kind = 0;
  long long N = 0, Ns = 0;  // particles on this rank, padded leading dimension
  int d = 0, nt = 0, ns = 0, items = 8;
  HostModel model;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void* x[2] = {nullptr, nullptr};  // ping-pong clouds; x[cur] is the current one
  int cur = 0;
  void* logw = nullptr;       // the log-weights of the last observed step
  void* logw_base = nullptr;  // the allocation: one buffer, or two for a sharded filter (by observed-step parity: a peer's walk
  size_t logw_stride = 0;     // over a run of repeated keys may read the previous step's while this rank writes the next)
  bool shared_device = false; // sharded: a peer rank lives on this device (one stream for all of them: k_publish behind each kernel)
  int32_t* anc = nullptr;  // GLOBAL particle indices of the last resampling (offspring slots of this rank)
  bool anc_valid = false, initialised = false;
  FilterScalars* sc = nullptr;
  SumTables tb = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  XchSlot* xch = nullptr;  // [MAXR] written by the peers (sharded filters)
  double* ubuf = nullptr;   // uniforms (stratified / multinomial, injected): one per GLOBAL output
  double* cdf = nullptr;    // N cumulative values (multinomial)
  double* scratch = nullptr;  // grow-only fp64 scratch (injected noise, read-back staging)
  size_t scratch_n = 0;
  void* ctab = nullptr;  // LGCP per-sub-step f coefficients (filter dtype)
  size_t ctab_n = 0;
  double *ll_steps = nullptr; int* ess_steps = nullptr; double* states = nullptr; size_t steps_cap = 0;
  std::vector<SeriesStep> series;
  std::vector<double> ctab_host;
  double t0_series = 0.0;
  double t_cur = 0.0;
  uint64_t seed = 0, stream_id = 0, epoch = 0;
  uint32_t key0 = 0, key1 = 0;
  uint32_t step_ctr = 0;
  // sharding: rank r of R owns the global slots [r*N, (r+1)*N)
  int rank = 0, world = 1;
  bool connected = false;
  unsigned long long obs_seq = 0, gstep = 0;  // monotone over the life of the handle (never reset)
  struct PeerPtrs { void* x[2]; int32_t* anc; void* logw; u128* tile_sum; double* tile_maxw; XchSlot* xch; };
  PeerPtrs peer[MAXR];
  std::vector<void*> ipc_opened;  // base pointers to close on destroy
  // single-launch series kernel (small clouds, cssm_series.cuh)
  u128* tile_q = nullptr;          // [nt] exact tile sums of w1^2
  SeriesCtl* series_ctl = nullptr;
  unsigned long long* anc64 = nullptr;  // k_series_small: tagged ancestor words (allocated on first use)
  void* logw2 = nullptr;                // k_series_small: second log-weight buffer
  u128* ser_tile_sum = nullptr;         // k_series_small: tile tables for its own tile size (256 * series_items particles)
  u128* ser_tile_q = nullptr;
  double* ser_tile_maxw = nullptr;
  int series_items = 0;                 // particles per thread of k_series_small: 1 or 2 (0: not eligible)
  void* recs = nullptr;            // per-observation records of the loaded series (filter dtype)
  size_t recs_cap = 0;             // bytes
  bool recs_valid = false;
  int series_mode = CSSM_SERIES_AUTO;
  int series_max_blocks = -1;      // co-resident blocks of k_series_small on this device (-1: not queried)
  int series_multi_blocks = 0;     // ... of k_series_multi
  long long series_multi_max = 1ll << 19;  // largest cloud the multi-tile series kernel is used for (CSSM_SERIES_MAX_N)
  bool series_use_multi = false;
  bool series_use_one = false;     // tiny cloud: the whole llFilter in one block (k_series_one)
  bool last_single_launch = false;
  int pdl = 1;  // programmatic dependent launch between the kernels of a step
  int flat_max_nt = 2048;  // clouds of at most this many tiles: K2 without atomics, K3 adds the tile sums itself (CSSM_FLAT_MAX_NT)
  int tie_first = 0;  // CSSM_TIE_FIRST instead of the reference's TreeMap rule (cssm_filter_set_tie_rule)
  int scan_fast = 1;  // k_scan_search tries the certified fp64 path first (cssm_filter_scan_mode; CSSM_K3_FAST=0)
  int scan_fast_min_nt = 1024;  // ... on clouds of more tiles than this (CSSM_K3_FAST=1: always).  A cloud of one or two waves of
                                // blocks lasts as long as its slowest block, and a tile the certified path hands to the exact
                                // path costs both: 2^20 particles 2.81e10 (exact only) vs 2.66e10, 2^21 equal, 2^22 3.83e10 vs 3.76e10
  // forecast cloud (cssm_forecast.cuh): d + 4 columns [x1 | gamma | eta | obs | obs2], filter dtype
  // path storage (FilterInterpolate): px = (paths_cap + 1) propagated clouds, panc = paths_cap ancestor vectors
  void* px = nullptr;
  int32_t* panc = nullptr;
  uint8_t* pres_dev = nullptr;
  long long paths_cap = 0, paths_len = -1;  // paths_len = recorded steps; -1: nothing recorded (no init since enable)
  std::vector<uint8_t> pres;                // per recorded step: did it resample?
  void* fc = nullptr;
  bool fc_valid = false;
  double fc_t = 0.0;
  uint32_t fc_ctr = 0;
  float last_ms = 0.f;
  // pinned staging of the small tables a PMMH iteration rebuilds (records, LGCP coefficients): asynchronous copies, the
  // buffer is reused only after the previous copy out of it has completed (pin_ev)
  void* pin = nullptr;
  size_t pin_cap = 0;
  cudaEvent_t pin_ev = nullptr;
  bool pin_pending = false;
  // per-kernel-class device timing (CUDA events on the launching stream), sampled every prof_stride steps
  int prof_stride = 0;
  std::vector<cudaEvent_t> prof_ev;  // pairs
  std::vector<int> prof_cls;
  double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_step = 0;
  long long last_launches = 0, launches = 0;
};

namespace {

void rekey(cssm_filter* f) {
  uint64_t k = splitmix64(f->seed ^ splitmix64(f->stream_id + 0x632BE59BD9B4E019ull) ^ splitmix64(f->epoch * 0xD1B54A32D192ED03ull + 1));
  f->key0 = (uint32_t)k;
  f->key1 = (uint32_t)(k >> 32);
}

int ensure_scratch(cssm_filter* f, size_t n) {
  if (n <= f->scratch_n) return CSSM_OK;
  if (f->scratch) cudaFree(f->scratch);
  f->scratch = nullptr; f->scratch_n = 0;
  CU(cudaMalloc(&f->scratch, n * sizeof(double)));
  f->scratch_n = n;
  return CSSM_OK;
}

// `bytes` of pinned host memory whose previous contents have left for the device
int pin_reserve(cssm_filter* f, size_t bytes, void** out) {
  if (f->pin_pending) {
    CU(cudaEventSynchronize(f->pin_ev));
    f->pin_pending = false;
  }
  if (bytes > f->pin_cap) {
    if (f->pin) cudaFreeHost(f->pin);
    f->pin = nullptr; f->pin_cap = 0;
    const size_t cap = std::max<size_t>(bytes, 1 << 16);
    CU(cudaMallocHost(&f->pin, cap));
    f->pin_cap = cap;
  }
  if (!f->pin_ev) CU(cudaEventCreateWithFlags(&f->pin_ev, cudaEventDisableTiming));
  *out = f->pin;
  return CSSM_OK;
}
// asynchronous copy of the staged bytes; no host wait
int pin_send(cssm_filter* f, void* dst, size_t bytes) {
  CU(cudaMemcpyAsync(dst, f->pin, bytes, cudaMemcpyHostToDevice, f->stream));
  CU(cudaEventRecord(f->pin_ev, f->stream));
  f->pin_pending = true;
  return CSSM_OK;
}

inline int nblk(long long n, int per) { return (int)((n + per - 1) / per); }

// kernel launch, optionally as a programmatic dependent launch: the kernel may be scheduled while
// the previous kernel of the stream drains; every such kernel starts with griddepcontrol.wait
template <typename... KArgs, typename... Args>
cudaError_t launch(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// the ranks of the filter as a kernel argument; `src` = which of the two clouds K1 gathers from
Peers make_peers(const cssm_filter* f, int src) {
  Peers pr;
  std::memset(&pr, 0, sizeof(pr));
  const bool sharded = f->world > 1 && f->connected;
  pr.R = sharded ? f->world : 1;
  pr.rank = sharded ? f->rank : 0;
  pr.Nl = f->N;
  pr.inv_nl = 1.0f / (float)f->N;
  for (int q = 0; q < pr.R; ++q) {
    if (q == pr.rank) {
      pr.x[q] = f->x[src]; pr.anc[q] = f->anc; pr.logw[q] = (const char*)f->logw_base + (size_t)(f->obs_seq & 1ull) * f->logw_stride;
      pr.tile_sum[q] = f->tb.tile_sum; pr.tile_maxw[q] = f->tb.tile_maxw; pr.xch[q] = f->xch;
    } else {
      const cssm_filter::PeerPtrs& pp = f->peer[q];
      pr.x[q] = pp.x[src]; pr.anc[q] = pp.anc; pr.logw[q] = (const char*)pp.logw + (size_t)(f->obs_seq & 1ull) * f->logw_stride;
      pr.tile_sum[q] = pp.tile_sum; pr.tile_maxw[q] = pp.tile_maxw; pr.xch[q] = pp.xch;
    }
  }
  return pr;
}

int need_connected(const cssm_filter* f) {
  if (f->world > 1 && !f->connected) return fail(CSSM_ERR_STATE, "sharded filter: cssm_filter_shard_connect has not been called");
  return CSSM_OK;
}

template <typename real>
int launch_init(cssm_filter* f, const double* zinj_dev, const double* x0) {
  StepArgs<real> a;
  std::memset(&a, 0, sizeof(a));
  a.d = f->d;
  int k = 0;
  for (const HostLeaf& L : f->model.leaves)
    for (int c = 0; c < L.dim; ++c, ++k) {
      a.S[k] = (real)std::sqrt(L.c0[c]);
      a.M[k] = (real)(x0 ? x0[k] : L.m0[c]);
    }
  const unsigned long long slot0 = (unsigned long long)f->rank * (unsigned long long)f->N;
  if (x0)
    k_fill_particles<real><<<nblk(f->N, 256), 256, 0, f->stream>>>(a, (real*)f->x[f->cur], f->N, f->Ns);
  else
    k_init_particles<real><<<nblk(f->N, 256), 256, 0, f->stream>>>(a, (real*)f->x[f->cur], zinj_dev, f->N, f->Ns, slot0, f->key0,
                                                                    f->key1, (uint32_t)f->epoch);
  f->launches++;
  CU(cudaGetLastError());
  return CSSM_OK;
}

int do_init(cssm_filter* f, double t0, const double* zinj_dev, const double* x0) {
  int rc = need_connected(f);
  if (rc) return rc;
  f->epoch++;
  rekey(f);
  f->step_ctr = 0;
  FilterScalars z;
  std::memset(&z, 0, sizeof(z));
  z.ess = (int)std::min<long long>(f->N * (long long)f->world, 2147483647LL);
  z.qb = 96;
  CU(cudaMemcpyAsync(f->sc, &z, sizeof(z), cudaMemcpyHostToDevice, f->stream));
  CU(cudaMemsetAsync(f->tb.super_sum, 0, (size_t)2 * f->ns * sizeof(u128), f->stream));
  CU(cudaMemsetAsync(f->tb.super_q, 0, (size_t)2 * f->ns * sizeof(u128), f->stream));
  CU(cudaMemsetAsync(f->tb.super_ticket, 0, (size_t)f->ns * sizeof(unsigned long long), f->stream));
  rc = (f->dtype == CSSM_F32) ? launch_init<float>(f, zinj_dev, x0) : launch_init<double>(f, zinj_dev, x0);
  if (rc) return rc;
  f->anc_valid = false;
  f->initialised = true;
  f->t_cur = t0;
  f->fc_valid = false;  // a forecast cloud does not outlive the filtering cloud it was drawn from
  if (f->paths_cap > 0) {  // FilterInterpolate: the initial cloud is element 0 of every path
    const size_t cloud = (size_t)f->d * f->Ns * (f->dtype == CSSM_F32 ? 4 : 8);
    CU(cudaMemcpyAsync(f->px, f->x[f->cur], cloud, cudaMemcpyDeviceToDevice, f->stream));
    f->paths_len = 0;
    f->pres.clear();
  }
  return CSSM_OK;
}

enum { CLS_PROPAGATE = 0, CLS_SUMS = 1, CLS_SEARCH = 2, CLS_MULTI = 3, CLS_INIT = 4, CLS_SERIES = 5 };

// event pair around one launch when profiling samples this step
struct ProfScope {
  cssm_filter* f;
  bool on;
  ProfScope(cssm_filter* f_, int cls, bool sampled) : f(f_), on(false) {
    if (!sampled) return;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    f->prof_ev.push_back(a);
    f->prof_ev.push_back(b);
    f->prof_cls.push_back(cls);
    cudaEventRecord(a, f->stream);
    on = true;
  }
  ~ProfScope() {
    if (on) cudaEventRecord(f->prof_ev.back(), f->stream);
  }
};

void prof_collect(cssm_filter* f) {
  for (size_t i = 0; i < f->prof_cls.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, f->prof_ev[2 * i], f->prof_ev[2 * i + 1]) == cudaSuccess) {
      f->prof_ms[f->prof_cls[i]] += ms;
      f->prof_n[f->prof_cls[i]] += 1;
    }
    cudaEventDestroy(f->prof_ev[2 * i]);
    cudaEventDestroy(f->prof_ev[2 * i + 1]);
  }
  f->prof_ev.clear();
  f->prof_cls.clear();
}

struct StepIO {
  const double* zinj = nullptr;   // device, [n_sub][d][N]
  const double* uarr = nullptr;   // device, one uniform per GLOBAL output (stratified / multinomial)
  int use_u_inj = 0;              // systematic uniform taken from sc->u_inj
  double* ll_steps = nullptr;     // device
  int* ess_steps = nullptr;       // device
  long long step_slot = 0;
};

// what one step needs to remember between its three launches
struct StepCtx {
  uint32_t step = 0;
  bool prof = false, observed = false;
  int parity = 0;
};

// ---- K1: gather + propagate + weight ------------------------------------------------------------
// Sharded filters whose ranks share a device (and therefore a stream): a one-warp kernel behind each producing kernel
// publishes its result to the peers.  One GPU per rank: block 0 of the consuming kernel does it (pub_here).
inline bool pub_here(const cssm_filter* f) { return f->world > 1 && !f->shared_device; }
int publish_after(cssm_filter* f, const Peers& pr, int what, int parity, unsigned long long value) {
  if (f->world == 1 || !f->shared_device) return CSSM_OK;
  k_publish<<<1, 32, 0, f->stream>>>(pr, f->sc, what, parity, value);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch k_publish: ") + cudaGetErrorString(e));
  return CSSM_OK;
}

template <typename real>
int step_phase1(cssm_filter* f, const StepHost& h, long long n_sub, const void* ctab, double delta, const StepIO& io, StepCtx& cx) {
  StepArgs<real> a;
  to_args<real>(f->model, h, a);
  const int32_t* anc = f->anc_valid ? f->anc : nullptr;
  cx.step = f->step_ctr++;
  cx.prof = f->prof_stride > 0 && (f->prof_step++ % f->prof_stride) == 0;
  cx.observed = h.has_obs != 0;
  cx.parity = (int)(f->obs_seq & 1ull);
  if (cx.observed) f->logw = (char*)f->logw_base + (size_t)cx.parity * f->logw_stride;
  const Peers pr = make_peers(f, f->cur);
  real* xdst = (real*)f->x[f->cur ^ 1];
  const unsigned long long slot0 = (unsigned long long)f->rank * (unsigned long long)f->N;
  static const bool shard_local = !(std::getenv("CSSM_SHARD_LOCAL") && std::atoi(std::getenv("CSSM_SHARD_LOCAL")) == 0);
  K1Ctl ctl{f->sc, cx.parity, f->obs_seq, f->gstep, shard_local ? 1 : 0, pub_here(f) ? 1 : 0};
  const bool pdl = f->pdl && !cx.prof && f->world == 1;  // sharded: stream order (block 0 of a kernel publishes the one before)
  cudaError_t e;
  {
    ProfScope ps_(f, CLS_PROPAGATE, cx.prof);
    if (f->model.obs_kind == CSSM_OBS_LGCP) {
      const int g = nblk(f->N, 256);
#define LGCP_CASE(DP)                                                                                                  \
  e = launch(k_lgcp_weight<real, DP>, g, 256, f->stream, pdl, a, pr, xdst, anc, (real*)f->logw, io.zinj, (const real*)ctab, \
             n_sub, (real)delta, f->N, f->Ns, slot0, f->key0, f->key1, cx.step, ctl)
      if (f->d <= 1) LGCP_CASE(1);
      else if (f->d <= 2) LGCP_CASE(2);
      else if (f->d <= 4) LGCP_CASE(4);
      else if (f->d <= 8) LGCP_CASE(8);
      else if (f->d <= 16) LGCP_CASE(16);
      else LGCP_CASE(32);
#undef LGCP_CASE
    } else {
      constexpr int PPT = VecOf<real>::PPT;
      const int g = nblk(f->N, 256 * PPT);
      // single-rank filters run instantiations without any sharding code (SH = false)
#define K1_CASE(DD)                                                                                                    \
  e = sharded ? launch(k_propagate_weight_sh<real, DD, PPT>, nblk(f->N, 256 * PPT * K1_SH_CHUNKS), 256, f->stream, pdl, a, pr, xdst, anc, (real*)f->logw, io.zinj, \
                       f->N, f->Ns, slot0, f->key0, f->key1, cx.step, ctl)                                              \
              : launch(k_propagate_weight<real, DD, PPT, false>, g, 256, f->stream, pdl, a, pr, xdst, anc, (real*)f->logw, io.zinj, \
                       f->N, f->Ns, slot0, f->key0, f->key1, cx.step, ctl)
      const bool sharded = pr.R > 1;
      static const int k1_ppt = std::getenv("CSSM_K1_PPT") ? std::atoi(std::getenv("CSSM_K1_PPT")) : 0;
      if (f->d == 7 && k1_ppt == 2 && sizeof(real) == 4) {
        e = launch(k_propagate_weight<float, 7, 2, false>, nblk(f->N, 256 * 2), 256, f->stream, pdl, *reinterpret_cast<StepArgs<float>*>(&a), pr,
                   (float*)xdst, anc, (float*)f->logw, io.zinj, f->N, f->Ns, slot0, f->key0, f->key1, cx.step, ctl);
      } else if (f->d == 1) K1_CASE(1);
      else if (f->d == 2) K1_CASE(2);
      else if (f->d == 7) K1_CASE(7);
      else K1_CASE(0);
#undef K1_CASE
    }
  }
  if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch K1: ") + cudaGetErrorString(e));
  f->launches++;
  {
    const int rcp = cx.observed ? publish_after(f, pr, 1, cx.parity, f->obs_seq) : publish_after(f, pr, 0, 0, f->gstep + 1);
    if (rcp) return rcp;
  }
  f->cur ^= 1;
  f->anc_valid = false;
  if (!cx.observed) f->gstep++;
  return CSSM_OK;
}

// the sum tables of one step: single-rank clouds of few tiles run flat (ns = 0, see SumTables)
inline SumTables step_tables(const cssm_filter* f) {
  SumTables tb = f->tb;
  if (f->world == 1 && f->nt <= f->flat_max_nt) tb.ns = 0;
  return tb;
}

// ---- K2: exact weight sums ----------------------------------------------------------------------
template <typename real>
int step_phase2(cssm_filter* f, StepCtx& cx) {
  if (!cx.observed) return CSSM_OK;
  const Peers pr = make_peers(f, f->cur);
  const bool pdl = f->pdl && !cx.prof && f->world == 1;  // sharded: stream order (block 0 of a kernel publishes the one before)
  cudaError_t e;
  {
    ProfScope ps_(f, CLS_SUMS, cx.prof);
#define K2_CASE(IT, SH)                                                                                               \
  e = launch(k_weight_sums<real, IT, SH>, f->nt, TILE_THREADS, f->stream, pdl, (const real*)f->logw, (const double*)nullptr, \
             f->N, f->sc, cx.parity, f->obs_seq, step_tables(f), pr, pub_here(f) ? 1 : 0)
    if (pr.R > 1) { if (f->items == 8) K2_CASE(8, true); else K2_CASE(2, true); }
    else { if (f->items == 8) K2_CASE(8, false); else K2_CASE(2, false); }
#undef K2_CASE
  }
  if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch K2: ") + cudaGetErrorString(e));
  f->launches++;
  return publish_after(f, pr, 2, cx.parity, f->obs_seq);
}

// K3 keeps two padded tiles of doubles in shared memory (37 KB per block with 2048-particle tiles):
// prefer the large shared-memory carveout so that four blocks fit an SM
// The attribute is per device and per kernel instantiation: one std::call_once per (real, device), so that handles on
// several devices of one process (ShardedGroup, one thread per GPU) each get it, and concurrent first calls are safe.
constexpr int MAX_DEVICES = 64;
template <typename real>
void k3_carveout_once(int device) {
  static std::once_flag once[MAX_DEVICES];
  if (device < 0 || device >= MAX_DEVICES) return;
  std::call_once(once[device], [] {
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_SYSTEMATIC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_STRATIFIED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_SYSTEMATIC>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_STRATIFIED>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_SYSTEMATIC, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_STRATIFIED, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_SYSTEMATIC, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_STRATIFIED, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_SYSTEMATIC, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 8, CSSM_RESAMPLE_STRATIFIED, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_SYSTEMATIC, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_scan_search<real, 2, CSSM_RESAMPLE_STRATIFIED, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaGetLastError();
  });
}

// ---- K3: CDF scan + ancestor search (+ multinomial draw) ----------------------------------------
template <typename real>
int step_phase3(cssm_filter* f, const StepIO& io, StepCtx& cx) {
  if (!cx.observed) return CSSM_OK;
  k3_carveout_once<real>(f->device);
  const Peers pr = make_peers(f, f->cur);
  const bool pdl = f->pdl && !cx.prof && f->world == 1;  // sharded: stream order (block 0 of a kernel publishes the one before)
  const bool multi = f->resample_kind == CSSM_RESAMPLE_MULTINOMIAL;
  K3Ctl ctl;
  ctl.parity = cx.parity; ctl.obs_seq = f->obs_seq; ctl.gstep = f->gstep;
  const long long Ng = f->N * (long long)pr.R;
  ctl.inv_n = ((Ng & (Ng - 1)) == 0) ? 1.0 / (double)Ng : 0.0;
  ctl.direct = 0; ctl.add_ll = 1; ctl.use_u_inj = io.use_u_inj; ctl.tie_first = f->tie_first; ctl.defer_ll = 0;
  ctl.fast_ok = f->scan_fast && f->nt > f->scan_fast_min_nt;
  ctl.anc64 = nullptr; ctl.anc_tag = 0; ctl.dbg = nullptr;
  ctl.key0 = f->key0; ctl.key1 = f->key1; ctl.step = cx.step;
  ctl.ll_steps = io.ll_steps; ctl.ess_steps = io.ess_steps; ctl.step_slot = io.step_slot;
  const bool strat = f->resample_kind == CSSM_RESAMPLE_STRATIFIED;
  const double* ua = multi ? (const double*)nullptr : io.uarr;
  double* cdf = multi ? f->cdf : (double*)nullptr;
  cudaError_t e;
  {
    ProfScope ps_(f, CLS_SEARCH, cx.prof);
    const SumTables tbs = step_tables(f);
#define K3_CASE(IT, KD)                                                                                                 \
  e = tbs.ns == 0 ? launch(k_scan_search<real, IT, KD, true, false>, f->nt, TILE_THREADS, f->stream, pdl, (const real*)f->logw, \
                           (const double*)nullptr, f->N, f->sc, tbs, pr, ctl, ua, cdf, pub_here(f) ? 1 : 0)                                  \
      : pr.R > 1  ? launch(k_scan_search<real, IT, KD, false, true>, f->nt, TILE_THREADS, f->stream, pdl, (const real*)f->logw, \
                           (const double*)nullptr, f->N, f->sc, tbs, pr, ctl, ua, cdf, pub_here(f) ? 1 : 0)                                  \
                  : launch(k_scan_search<real, IT, KD, false, false>, f->nt, TILE_THREADS, f->stream, pdl, (const real*)f->logw, \
                           (const double*)nullptr, f->N, f->sc, tbs, pr, ctl, ua, cdf, pub_here(f) ? 1 : 0)
    if (f->items == 8) { if (strat) K3_CASE(8, CSSM_RESAMPLE_STRATIFIED); else K3_CASE(8, CSSM_RESAMPLE_SYSTEMATIC); }
    else { if (strat) K3_CASE(2, CSSM_RESAMPLE_STRATIFIED); else K3_CASE(2, CSSM_RESAMPLE_SYSTEMATIC); }
#undef K3_CASE
  }
  if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch K3: ") + cudaGetErrorString(e));
  f->launches++;
  if (multi) {
    ProfScope ps_(f, CLS_MULTI, cx.prof);
    e = launch(k_multinomial_search, nblk(f->N, 256), 256, f->stream, pdl, (const double*)f->cdf, f->N, io.uarr, f->key0, f->key1,
               cx.step, f->anc, &f->sc->flags);
    if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch K4: ") + cudaGetErrorString(e));
    f->launches++;
  }
  f->anc_valid = true;
  f->obs_seq++;
  f->gstep++;
  return publish_after(f, pr, 0, 0, f->gstep);  // "resampling done": the number of completed steps
}

// one stepFilter on the device; no host synchronisation
template <typename real>
int launch_step(cssm_filter* f, const StepHost& h, long long n_sub, const void* ctab, double delta, const StepIO& io) {
  StepCtx cx;
  int rc = step_phase1<real>(f, h, n_sub, ctab, delta, io, cx);
  if (rc) return rc;
  rc = step_phase2<real>(f, cx);
  if (rc) return rc;
  return step_phase3<real>(f, io, cx);
}

int step_consts(cssm_filter* f, double t_prev, double t, int has_obs, double y, StepHost& h, long long& n_sub, double& delta) {
  const HostModel& m = f->model;
  double dt = t - t_prev;
  n_sub = 1;
  delta = 0.0;
  if (m.obs_kind == CSSM_OBS_LGCP) {
    delta = std::pow(10, -m.lgcp_precision);
    n_sub = (dt == 0) ? 0 : (long long)(int)std::ceil(dt / delta);  // model/ParticleFilter.scala:190
    if (n_sub >= (1 << 22)) return fail(CSSM_ERR_UNSUPPORTED, "LGCP: more than 2^22 sub-steps in one increment");
    // the Philox call counter of a step shares its word with the 8-bit purpose tag (RNG_STEP | call): n_sub * d / PER_CALL
    // calls must stay below 2^24 or the stream would run into the resampling / sample-one streams of the same step
    if (n_sub * (long long)m.d / (f->dtype == CSSM_F32 ? 4 : 2) >= (1ll << 24))
      return fail(CSSM_ERR_UNSUPPORTED, "LGCP: sub-steps x latent dimension exceeds the 2^24 Philox calls of one step");
    transition_consts(m, delta, h);
    has_obs = 1;  // every datum is an event; FilterLgcp always weights and resamples (:217-225)
  } else {
    transition_consts(m, dt, h);
  }
  f_coeffs(m, t, h.C);
  obs_consts(m, has_obs, y, h);
  h.has_obs = has_obs;
  return CSSM_OK;
}

bool has_seasonal(const HostModel& m) {
  for (const HostLeaf& L : m.leaves)
    if (L.f_kind == CSSM_F_SEASONAL) return true;
  return false;
}

// LGCP with a time-dependent f: coefficients at the sub-step times t_i = t + i*delta (accumulated,
// the stream starts at the observation time, model/ParticleFilter.scala:194,215)
void lgcp_ctab_host(const HostModel& m, double t, long long n_sub, double delta, std::vector<double>& out) {
  double time = t;
  size_t o = out.size();
  out.resize(o + (size_t)n_sub * m.d);
  for (long long s = 0; s < n_sub; ++s) {
    time = time + delta;
    f_coeffs(m, time, out.data() + o + (size_t)s * m.d);
  }
}

int upload_ctab(cssm_filter* f, const std::vector<double>& host) {
  if (host.empty()) return CSSM_OK;
  size_t esz = (f->dtype == CSSM_F32) ? 4 : 8;
  if (host.size() > f->ctab_n) {
    if (f->ctab) cudaFree(f->ctab);
    f->ctab = nullptr; f->ctab_n = 0;
    CU(cudaMalloc(&f->ctab, host.size() * esz));
    f->ctab_n = host.size();
  }
  void* stage;
  int rc = pin_reserve(f, host.size() * esz, &stage);
  if (rc) return rc;
  if (f->dtype == CSSM_F32) {
    float* o = (float*)stage;
    for (size_t i = 0; i < host.size(); ++i) o[i] = (float)host[i];
  } else {
    std::memcpy(stage, host.data(), host.size() * 8);
  }
  return pin_send(f, f->ctab, host.size() * esz);
}

// constants (+ LGCP coefficient table) of one step taken outside a loaded series
int prepare_one_step(cssm_filter* f, double t, int has_obs, double y, StepHost& h, long long& n_sub, double& delta, const void*& ctab) {
  int rc = step_consts(f, f->t_cur, t, has_obs, y, h, n_sub, delta);
  if (rc) return rc;
  ctab = nullptr;
  if (f->model.obs_kind == CSSM_OBS_LGCP && has_seasonal(f->model) && n_sub > 0) {
    std::vector<double> host;
    lgcp_ctab_host(f->model, t, n_sub, delta, host);
    rc = upload_ctab(f, host);
    if (rc) return rc;
    ctab = f->ctab;
  }
  return CSSM_OK;
}

int run_one_step(cssm_filter* f, double t, int has_obs, double y, const StepIO& io) {
  StepHost h;
  long long n_sub;
  double delta;
  const void* ctab;
  int rc = prepare_one_step(f, t, has_obs, y, h, n_sub, delta, ctab);
  if (rc) return rc;
  if (f->paths_cap > 0 && f->paths_len >= f->paths_cap) return fail(CSSM_ERR_STATE, "path storage is full (cssm_filter_paths_enable)");
  rc = (f->dtype == CSSM_F32) ? launch_step<float>(f, h, n_sub, ctab, delta, io) : launch_step<double>(f, h, n_sub, ctab, delta, io);
  if (rc) return rc;
  f->t_cur = t;
  f->fc_valid = false;
  if (f->paths_cap > 0 && f->paths_len >= 0) {  // FilterInterpolate: keep the propagated cloud and the ancestors of this step
    const size_t cloud = (size_t)f->d * f->Ns * (f->dtype == CSSM_F32 ? 4 : 8);
    const bool resampled = f->anc_valid;
    CU(cudaMemcpyAsync((char*)f->px + (size_t)(f->paths_len + 1) * cloud, f->x[f->cur], cloud, cudaMemcpyDeviceToDevice, f->stream));
    if (resampled)
      CU(cudaMemcpyAsync(f->panc + (size_t)f->paths_len * f->N, f->anc, (size_t)f->N * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                         f->stream));
    f->pres.push_back(resampled ? 1 : 0);
    f->paths_len++;
  }
  return CSSM_OK;
}

int read_ll(cssm_filter* f, double* ll, int32_t* ess) {
  FilterScalars s;
  CU(cudaMemcpyAsync(&s, f->sc, sizeof(s), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  if (!f->prof_cls.empty()) prof_collect(f);
  if (ll) *ll = s.ll;
  if (ess) *ess = s.ess;
  if (s.flags & FLAG_COMM_TIMEOUT) return fail(CSSM_ERR_COMM, "sharded filter: a peer rank did not arrive within 4 s");
  return CSSM_OK;
}

int enter(const cssm_filter* f) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  CU(cudaSetDevice(f->device));
  return CSSM_OK;
}

int ensure_steps_cap(cssm_filter* f, size_t T) {
  if (T + 1 <= f->steps_cap) return CSSM_OK;
  if (f->ll_steps) cudaFree(f->ll_steps);
  if (f->ess_steps) cudaFree(f->ess_steps);
  if (f->states) cudaFree(f->states);
  f->ll_steps = nullptr; f->ess_steps = nullptr; f->states = nullptr; f->steps_cap = 0;
  CU(cudaMalloc(&f->ll_steps, (T + 1) * sizeof(double)));
  CU(cudaMalloc(&f->ess_steps, (T + 1) * sizeof(int)));
  CU(cudaMalloc(&f->states, (T + 1) * (size_t)f->d * sizeof(double)));
  f->steps_cap = T + 1;
  return CSSM_OK;
}

template <typename real>
void launch_sample_one(cssm_filter* f, double* out_dev, uint32_t tag) {
  const Peers pr = make_peers(f, f->cur);
  k_sample_one<real><<<1, 32, 0, f->stream>>>(pr, f->anc_valid ? f->anc : nullptr, out_dev, f->d, f->N, f->Ns, f->key0, f->key1, tag);
  f->launches++;
}

template <typename real>
void launch_gather(cssm_filter* f, const int32_t* anc, double* out_dev) {
  const Peers pr = make_peers(f, f->cur);
  k_gather<real, double><<<nblk(f->N, 256), 256, 0, f->stream>>>(pr, anc, out_dev, f->d, f->N, f->Ns, f->N);
}

// ---- single-launch series kernels: instantiated in their own translation unit (cssm_series.cu) -------------------
void* series_kernel(const cssm_filter* f, int items) { return cssm::series_small_kernel(f->dtype, items, f->d, f->resample_kind); }
void* series_multi_kernel(const cssm_filter* f) { return cssm::series_multi_kernel(f->dtype, f->items, f->d, f->resample_kind); }
int resident_blocks(const cssm_filter* f, void* kern) {
  int per_sm = 0, sms = 0, coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, f->device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE_THREADS, 0) != cudaSuccess) per_sm = 0;
  cudaGetLastError();
  return coop ? per_sm * sms : 0;
}

// may this filter's loaded series run as one launch?  One 512-particle tile per block, every block resident.
bool series_eligible(cssm_filter* f, bool sample_states) {
  if (f->series_mode == CSSM_SERIES_THREE_LAUNCH || sample_states || f->world > 1) return false;
  if (f->model.obs_kind == CSSM_OBS_LGCP || f->resample_kind == CSSM_RESAMPLE_MULTINOMIAL) return false;
  if (f->series.empty() || f->series.size() > 0x7fffffffull) return false;
  if (f->series_max_blocks < 0) {
    f->series_max_blocks = std::min(resident_blocks(f, series_kernel(f, 2)), SERIES_SMALL_MAX_TILES);
    f->series_multi_blocks = resident_blocks(f, series_multi_kernel(f));
  }
  f->series_use_multi = false;
  f->series_use_one = false;
  static const bool no_one = std::getenv("CSSM_SERIES_ONE") != nullptr && std::atoi(std::getenv("CSSM_SERIES_ONE")) == 0;
  if (f->N <= SERIES_ONE_MAX_N && !no_one) {  // the reference's own particle counts: one block, no grid-wide exchange at all
    f->series_use_one = true;
    return true;
  }
  // one 512-particle tile per block, weights stay in registers
  f->series_items = 0;
  if (nblk(f->N, 2 * TILE_THREADS) + 1 <= f->series_max_blocks) f->series_items = 2;  // one block per tile + the accountant
  if (f->series_items != 0) return true;
  // several tiles per block: pays while a stage is short against the launch gaps it removes
  if (f->series_multi_blocks > 0 && (f->N <= f->series_multi_max || f->series_mode == CSSM_SERIES_SINGLE_LAUNCH)) {
    f->series_use_multi = true;
    return true;
  }
  return false;
}

// the per-observation records of the loaded series: [A D S C | y k0 k1 k2 k3 has_obs 0 0] in the filter dtype
template <typename real>
int upload_recs(cssm_filter* f) {
  const size_t T = f->series.size(), len = (size_t)4 * f->d + SERIES_REC_EXTRA;
  const size_t bytes = T * len * sizeof(real);
  void* stage;
  int rc = pin_reserve(f, bytes, &stage);
  if (rc) return rc;
  real* host = (real*)stage;
  for (size_t s = 0; s < T; ++s) {
    StepArgs<real> a;
    to_args<real>(f->model, f->series[s].h, a);
    real* r = host + s * len;
    for (int k = 0; k < f->d; ++k) {
      r[k] = a.A[k]; r[f->d + k] = a.D[k]; r[2 * f->d + k] = a.S[k]; r[3 * f->d + k] = a.C[k];
    }
    real* e = r + 4 * f->d;
    e[0] = a.y; e[1] = a.k0; e[2] = a.k1; e[3] = a.k2; e[4] = a.k3; e[5] = a.has_obs ? (real)1 : (real)0; e[6] = e[7] = (real)0;
  }
  if (bytes > f->recs_cap) {
    if (f->recs) cudaFree(f->recs);
    f->recs = nullptr; f->recs_cap = 0;
    CU(cudaMalloc(&f->recs, bytes));
    f->recs_cap = bytes;
  }
  rc = pin_send(f, f->recs, bytes);
  if (rc) return rc;
  f->recs_valid = true;
  return CSSM_OK;
}

int run_series_single_launch(cssm_filter* f) {
  const size_t T = f->series.size();
  int rc = CSSM_OK;
  if (!f->recs_valid) rc = (f->dtype == CSSM_F32) ? upload_recs<float>(f) : upload_recs<double>(f);
  if (rc) return rc;
  CU(cudaEventRecord(f->ev0, f->stream));
  rc = do_init(f, f->t0_series, nullptr, nullptr);
  if (rc) return rc;
  CU(cudaMemsetAsync(f->series_ctl, 0, sizeof(SeriesCtl), f->stream));
  SeriesArgs sa;
  std::memset(&sa, 0, sizeof(sa));
  if (!f->series_use_multi && !f->series_use_one) {  // one tile per block: tagged ancestors instead of a third grid barrier (cssm_series.cuh)
    const size_t esz = (f->dtype == CSSM_F32) ? 4 : 8;
    if (f->anc64 == nullptr) {
      if (cudaMalloc((void**)&f->anc64, (size_t)f->Ns * sizeof(unsigned long long)) != cudaSuccess ||
          cudaMalloc(&f->logw2, (size_t)(f->Ns + TILE_THREADS * f->items) * esz) != cudaSuccess)
        return fail(CSSM_ERR_NOMEM, "cudaMalloc: series kernel buffers");
      CU(cudaMemsetAsync(f->logw2, 0, (size_t)(f->Ns + TILE_THREADS * f->items) * esz, f->stream));
      const size_t snt = (size_t)nblk(f->N, TILE_THREADS) + 1;
      if (cudaMalloc((void**)&f->ser_tile_sum, snt * sizeof(u128)) != cudaSuccess || cudaMalloc((void**)&f->ser_tile_q, snt * sizeof(u128)) != cudaSuccess ||
          cudaMalloc((void**)&f->ser_tile_maxw, snt * sizeof(double)) != cudaSuccess)
        return fail(CSSM_ERR_NOMEM, "cudaMalloc: series kernel tables");
    }
    CU(cudaMemsetAsync(f->anc64, 0, (size_t)f->Ns * sizeof(unsigned long long), f->stream));  // tags of earlier launches
    sa.anc64 = f->anc64;
    sa.logw2 = f->logw2;
  }

  sa.x[0] = f->x[f->cur]; sa.x[1] = f->x[f->cur ^ 1];
  sa.logw = f->logw; sa.anc = f->anc; sa.sc = f->sc;
  sa.tile_sum = f->tb.tile_sum; sa.tile_q = f->tile_q; sa.tile_maxw = f->tb.tile_maxw;
  sa.nt = f->nt;
  if (!f->series_use_multi && !f->series_use_one) {
    sa.tile_sum = f->ser_tile_sum; sa.tile_q = f->ser_tile_q; sa.tile_maxw = f->ser_tile_maxw;
    sa.nt = nblk(f->N, TILE_THREADS * f->series_items);
  }
  sa.ctl = f->series_ctl; sa.recs = f->recs; sa.ll_steps = f->ll_steps; sa.ess_steps = f->ess_steps;
  sa.N = f->N; sa.Ns = f->Ns; sa.T = (int)T; sa.d = f->d; sa.obs_kind = f->model.obs_kind;
  sa.key0 = f->key0; sa.key1 = f->key1; sa.step0 = f->step_ctr;
  sa.inv_n = ((f->N & (f->N - 1)) == 0) ? 1.0 / (double)f->N : 0.0;
  sa.tie_first = f->tie_first;
  sa.pr[0] = make_peers(f, f->cur);
  sa.pr[1] = make_peers(f, f->cur ^ 1);
  static const bool debug_stamps = std::getenv("CSSM_SERIES_DEBUG") != nullptr;
  if (f->series_use_one) {  // per-step records of the exact sums: ll / ESS of all steps are evaluated at the end of the launch
    int rc2 = ensure_scratch(f, T * 6);
    if (rc2) return rc2;
    CU(cudaMemsetAsync(f->scratch, 0, T * 6 * sizeof(unsigned long long), f->stream));
    sa.ll_rec = (unsigned long long*)f->scratch;
  }
  if (debug_stamps && !f->series_use_multi && !f->series_use_one) {
    int rc2 = ensure_scratch(f, (size_t)16 * 512);
    if (rc2) return rc2;
    CU(cudaMemsetAsync(f->scratch, 0, (size_t)16 * 512 * 8, f->stream));
    sa.dbg = (unsigned long long*)f->scratch;
  }
  void* args[] = {&sa};
  const bool prof = f->prof_stride > 0;
  cudaError_t e;
  {
    ProfScope ps_(f, CLS_SERIES, prof);
    if (f->series_use_one)
      e = cudaLaunchKernel(cssm::series_one_kernel(f->dtype, f->d, f->resample_kind), dim3(1), dim3(TILE_THREADS), args, 0, f->stream);
    else if (f->series_use_multi)
      e = cudaLaunchCooperativeKernel(series_multi_kernel(f), dim3((unsigned)std::min(f->nt, f->series_multi_blocks)), dim3(TILE_THREADS),
                                      args, 0, f->stream);
    else
      e = cudaLaunchCooperativeKernel(series_kernel(f, f->series_items), dim3((unsigned)sa.nt + 1u), dim3(TILE_THREADS), args, 0, f->stream);  // + the accountant block
  }
  if (e != cudaSuccess) return fail(CSSM_ERR_CUDA, std::string("launch series kernel: ") + cudaGetErrorString(e));
  f->launches++;
  // host mirror of what the kernel did: T steps, the cloud flipped T times, ancestors valid iff the last step was observed
  f->step_ctr += (uint32_t)T;
  f->cur ^= (int)(T & 1);
  f->anc_valid = f->series.back().h.has_obs != 0;
  f->t_cur = f->series.back().t;
  CU(cudaEventRecord(f->ev1, f->stream));
  if (sa.dbg != nullptr && sa.nt <= 512) {  // per stage: thread 0's cycles per step, min / mean / max over the blocks (and who was slowest)
    static unsigned long long c[16 * 512];
    CU(cudaMemcpyAsync(c, f->scratch, sizeof(c), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    const char* name[16] = {"P1", "B1", "P2", "B2", "P3", "anc-wait", "head", "-", "P3:sums", "P3:finish", "P3:counts", "P3:ties", "P3:expand", "P3:sync", "-", "-"};
    const int order[13] = {6, 5, 0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 13};
    std::fprintf(stderr, "series kernel, cycles per step (T=%zu, blocks %d): ", T, sa.nt);
    for (int k : order) {
      double mn = 1e30, mx = 0, sum = 0;
      int who = 0;
      for (int b = 0; b < sa.nt; ++b) {
        const double v = c[(size_t)k * sa.nt + b] / (double)T;
        sum += v;
        if (v < mn) mn = v;
        if (v > mx) { mx = v; who = b; }
      }
      std::fprintf(stderr, "%s %.0f/%.0f/%.0f(b%d) ", name[k], mn, sum / sa.nt, mx, who);
    }
    std::fprintf(stderr, "\n");
  }
  return CSSM_OK;
}

// init + T steps on the loaded series; optionally one sampled particle per time (filter, :152-158)
int run_series_impl(cssm_filter* f, bool sample_states);
int run_series(cssm_filter* f, bool sample_states) {
  const int rc = run_series_impl(f, sample_states);
  f->fc_valid = false;
  if (f->paths_cap > 0) {  // the whole-series calls do not record paths: what was recorded no longer describes the cloud
    f->paths_len = -1;
    f->pres.clear();
  }
  return rc;
}
int run_series_impl(cssm_filter* f, bool sample_states) {
  const size_t T = f->series.size();
  int rc = ensure_steps_cap(f, T);
  if (rc) return rc;
  if (sample_states && f->world > 1) return fail(CSSM_ERR_UNSUPPORTED, "filter(): per-time sampled states are not available on a sharded filter");
  f->launches = 0;
  f->last_single_launch = series_eligible(f, sample_states);
  if (f->last_single_launch) return run_series_single_launch(f);
  if (f->series_mode == CSSM_SERIES_SINGLE_LAUNCH)
    return fail(CSSM_ERR_UNSUPPORTED, "the single-launch series kernel does not cover this filter (cloud too large for one resident "
                                      "grid, LGCP, multinomial, sharded, or per-time sampled states)");
  CU(cudaEventRecord(f->ev0, f->stream));
  rc = do_init(f, f->t0_series, nullptr, nullptr);
  if (rc) return rc;
  if (sample_states) {
    if (f->dtype == CSSM_F32) launch_sample_one<float>(f, f->states, 0x80000000u);
    else launch_sample_one<double>(f, f->states, 0x80000000u);
  }
  for (size_t s = 0; s < T; ++s) {
    const SeriesStep& st = f->series[s];
    StepIO io;
    io.ll_steps = f->ll_steps;
    io.ess_steps = f->ess_steps;
    io.step_slot = (long long)s;
    const void* ctab = nullptr;
    if (f->ctab && st.n_sub > 0 && !f->ctab_host.empty())
      ctab = (const char*)f->ctab + st.ctab_off * ((f->dtype == CSSM_F32) ? 4 : 8);
    double delta = (f->model.obs_kind == CSSM_OBS_LGCP) ? std::pow(10, -f->model.lgcp_precision) : 0.0;
    rc = (f->dtype == CSSM_F32) ? launch_step<float>(f, st.h, st.n_sub, ctab, delta, io)
                                : launch_step<double>(f, st.h, st.n_sub, ctab, delta, io);
    if (rc) return rc;
    f->t_cur = st.t;
    if (sample_states) {
      if (f->dtype == CSSM_F32) launch_sample_one<float>(f, f->states + (s + 1) * f->d, 0x80000001u + (uint32_t)s);
      else launch_sample_one<double>(f, f->states + (s + 1) * f->d, 0x80000001u + (uint32_t)s);
    }
  }
  CU(cudaEventRecord(f->ev1, f->stream));
  return CSSM_OK;
}

// what one rank tells the others about itself (cssm_filter_shard_export)
struct ShardBlob {
  uint64_t magic;
  int32_t rank, world, device, dtype, nt, d;
  int64_t N, Ns;
  int64_t pid;
  void* ptr[7];  // x0, x1, anc, logw, tile_sum, tile_maxw, xch
  cudaIpcMemHandle_t h[7];
};
static_assert(sizeof(ShardBlob) <= CSSM_SHARD_BLOB_BYTES, "blob size");
const uint64_t BLOB_MAGIC = 0x4353534d53484152ull;  // "CSSMSHAR"

int create_impl(const cssm_model_desc_t* model, int64_t n_particles, int resample_kind, int dtype, int device, uint64_t seed,
                uint64_t stream_id, int rank, int world, cssm_filter_t** out) {
  if (!out) return fail(CSSM_ERR_INVALID, "null output handle");
  *out = nullptr;
  if (world < 1 || world > MAXR || rank < 0 || rank >= world) return fail(CSSM_ERR_INVALID, "rank/world out of range (at most 8 ranks)");
  if (n_particles <= 0 || n_particles * (int64_t)world > 2147483647LL) return fail(CSSM_ERR_INVALID, "total particle count must be in [1, 2^31-1]");
  if (resample_kind < 0 || resample_kind > 2) return fail(CSSM_ERR_INVALID, "unknown resample_kind");
  if (world > 1 && resample_kind == CSSM_RESAMPLE_MULTINOMIAL) return fail(CSSM_ERR_UNSUPPORTED, "multinomial resampling is not available on a sharded filter");
  if (dtype != CSSM_F32 && dtype != CSSM_F64) return fail(CSSM_ERR_INVALID, "unknown dtype");
  HostModel hm;
  int rc = copy_model(model, hm);
  if (rc) return rc;
  int ndev = 0;
  rc = cssm_device_count(&ndev);
  if (rc) return rc;
  if (ndev == 0) return fail(CSSM_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(CSSM_ERR_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  cssm_filter* f = new cssm_filter();
  f->device = device; f->dtype = dtype; f->resample_kind = resample_kind;
  f->N = n_particles; f->Ns = (n_particles + 63) / 64 * 64;
  // experiment knob: extra elements of leading dimension (a power-of-two stride between the coordinate arrays of a cloud
  // puts the d streams of a thread on the same address bits)
  if (const char* e = std::getenv("CSSM_NS_PAD")) f->Ns += (std::atoll(e) + 63) / 64 * 64;
  f->model = hm; f->d = hm.d;
  f->rank = rank; f->world = world;
  // small clouds: 512-particle tiles so that the weight passes still fill the machine
  f->items = (f->N <= (1 << 18)) ? 2 : 8;
  if (const char* e = std::getenv("CSSM_TILE_ITEMS")) { int v = std::atoi(e); if (v == 2 || v == 8) f->items = v; }
  if (const char* e = std::getenv("CSSM_PDL")) f->pdl = std::atoi(e) != 0;
  if (const char* e = std::getenv("CSSM_SERIES_MAX_N")) f->series_multi_max = std::atoll(e);
  if (const char* e = std::getenv("CSSM_SERIES_KERNEL")) f->series_mode = std::atoi(e) ? CSSM_SERIES_AUTO : CSSM_SERIES_THREE_LAUNCH;
  const int tile = TILE_THREADS * f->items;
  f->nt = nblk(f->N, tile);
  f->ns = nblk(f->nt, SUPER);
  f->seed = seed; f->stream_id = stream_id;
  const size_t esz = (dtype == CSSM_F32) ? 4 : 8;
#define ALLOC(ptr, bytes)                                                                        \
  do {                                                                                           \
    cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                        \
    if (e_ != cudaSuccess) {                                                                     \
      cssm_filter_destroy(f);                                                                    \
      return fail(CSSM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e_));         \
    }                                                                                            \
  } while (0)
  ALLOC(f->x[0], (size_t)f->d * f->Ns * esz);
  ALLOC(f->x[1], (size_t)f->d * f->Ns * esz);
  f->logw_stride = (world > 1) ? (((size_t)(f->Ns + tile) * esz + 255) / 256) * 256 : 0;
  ALLOC(f->logw_base, (world > 1) ? 2 * f->logw_stride : (size_t)(f->Ns + tile) * esz);
  f->logw = f->logw_base;
  ALLOC(f->anc, (size_t)f->Ns * sizeof(int32_t));
  ALLOC(f->sc, sizeof(FilterScalars));
  ALLOC(f->xch, sizeof(XchSlot) * MAXR);
  ALLOC(f->tb.tile_sum, (size_t)(f->nt + 1) * sizeof(u128));
  ALLOC(f->tb.tile_maxw, (size_t)(f->nt + 1) * sizeof(double));
  ALLOC(f->tb.super_sum, (size_t)2 * f->ns * sizeof(u128));
  ALLOC(f->tb.super_q, (size_t)2 * f->ns * sizeof(u128));
  ALLOC(f->tb.super_ticket, (size_t)f->ns * sizeof(unsigned long long));
  if (resample_kind == CSSM_RESAMPLE_MULTINOMIAL) ALLOC(f->cdf, (size_t)f->Ns * sizeof(double));
  ALLOC(f->tile_q, (size_t)(f->nt + 1) * sizeof(u128));
  ALLOC(f->series_ctl, sizeof(SeriesCtl));
#undef ALLOC
  f->tb.nt = f->nt; f->tb.ns = f->ns;
  f->tb.tile_q = f->tile_q;
  if (const char* e = std::getenv("CSSM_FLAT_MAX_NT")) f->flat_max_nt = std::atoi(e);
  if (const char* e = std::getenv("CSSM_K3_FAST")) { f->scan_fast = std::atoi(e) != 0; if (f->scan_fast) f->scan_fast_min_nt = 0; }
  {
    const cudaError_t em[] = {cudaMemset(f->x[0], 0, (size_t)f->d * f->Ns * esz),
                              cudaMemset(f->x[1], 0, (size_t)f->d * f->Ns * esz),
                              cudaMemset(f->logw_base, 0, (world > 1) ? 2 * f->logw_stride : (size_t)(f->Ns + tile) * esz),
                              cudaMemset(f->anc, 0, (size_t)f->Ns * sizeof(int32_t)),
                              cudaMemset(f->sc, 0, sizeof(FilterScalars)),
                              cudaMemset(f->xch, 0, sizeof(XchSlot) * MAXR),
                              cudaMemset(f->tb.super_sum, 0, (size_t)2 * f->ns * sizeof(u128)),
                              cudaMemset(f->tb.super_q, 0, (size_t)2 * f->ns * sizeof(u128)),
                              cudaMemset(f->tb.super_ticket, 0, (size_t)f->ns * sizeof(unsigned long long)),
                              cudaMemset(f->series_ctl, 0, sizeof(SeriesCtl))};
    for (cudaError_t e_ : em)
      if (e_ != cudaSuccess) {
        cssm_filter_destroy(f);
        return fail(CSSM_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e_));
      }
  }
  if (cudaStreamCreateWithFlags(&f->own_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&f->ev0) != cudaSuccess ||
      cudaEventCreate(&f->ev1) != cudaSuccess) {
    cssm_filter_destroy(f);
    return fail(CSSM_ERR_CUDA, "stream/event creation failed");
  }
  f->stream = f->own_stream;
  CU(cudaDeviceSynchronize());
  *out = f;
  return CSSM_OK;
}

}  // namespace

// ---- forecast (cssm_forecast.cuh) ---------------------------------------------------------------
namespace {
template <typename real>
int forecast_impl(cssm_filter* f, double t, double interval, int chain, bool summarise, std::vector<double>& host) {
  const int d = f->d, cols = d + 4;
  const long long n = f->N;
  if (!f->fc) CU(cudaMalloc(&f->fc, (size_t)cols * f->Ns * sizeof(real)));
  const double t_from = chain ? f->fc_t : f->t_cur;
  StepHost h;
  std::memset(&h, 0, sizeof(h));
  transition_consts(f->model, t - t_from, h);
  f_coeffs(f->model, t, h.C);
  StepArgs<real> a;
  to_args<real>(f->model, h, a);
  // getCredibleInterval ranks for the state, getOrderStatistic ranks for eta and the observations (see cssm_filter_intervals).
  // Everything that can fail is checked BEFORE the forecast cloud is advanced: a rejected call leaves it (and the Philox
  // counter of the forecast stream) untouched, so a retry does not double-step a chained forecast.
  const long long idx = (long long)std::floor(interval * (double)n);
  const long long lo_s = n - idx - 1, hi_s = idx - 1, lo_e = n - idx, hi_e = idx;
  if (summarise) {
    if (lo_s < 0 || lo_s >= n || hi_s < 0 || hi_s >= n || lo_e < 0 || lo_e >= n || hi_e < 0 || hi_e >= n)
      return fail(CSSM_ERR_INVALID, "forecast: the order-statistic index is outside the cloud (the reference throws IndexOutOfBounds here)");
    // scratch: mean[cols] | out[2*cols] | SelState[2*cols] | hist[cols*512 u32]
    int rc = ensure_scratch(f, (size_t)cols + 2 * cols + 2 * (2 * cols) + (size_t)cols * 256 + 8);
    if (rc) return rc;
  }
  const Peers pr = make_peers(f, f->cur);
  const int32_t* anc = f->anc_valid ? f->anc : nullptr;
  ObsDraw od{f->model.obs_kind, f->model.obs_df, f->model.has_scale, f->model.scale};
  k_forecast<real><<<nblk(n, 256), 256, 0, f->stream>>>(a, pr, anc, (real*)f->fc, chain, od, n, f->Ns, f->key0, f->key1, f->fc_ctr);
  f->launches++;
  CU(cudaGetLastError());
  f->fc_ctr++;
  f->fc_valid = true;
  f->fc_t = t;
  if (!summarise) return CSSM_OK;
  double* mean_dev = f->scratch;
  double* out_dev = mean_dev + cols;
  SelState* sel_dev = reinterpret_cast<SelState*>(out_dev + 2 * cols);
  unsigned* hist_dev = reinterpret_cast<unsigned*>(sel_dev + 2 * cols);
  std::vector<SelState> sel((size_t)2 * cols);
  for (int c = 0; c < cols; ++c) {
    sel[2 * c] = SelState{0ull, c < d ? lo_s : lo_e};
    sel[2 * c + 1] = SelState{0ull, c < d ? hi_s : hi_e};
  }
  CU(cudaMemsetAsync(mean_dev, 0, (size_t)cols * sizeof(double), f->stream));
  CU(cudaMemsetAsync(hist_dev, 0, (size_t)cols * 512 * sizeof(unsigned), f->stream));
  CU(cudaMemcpyAsync(sel_dev, sel.data(), sel.size() * sizeof(SelState), cudaMemcpyHostToDevice, f->stream));
  Peers pc;  // the forecast cloud as a plain single-rank cloud of `cols` coordinates
  std::memset(&pc, 0, sizeof(pc));
  pc.R = 1; pc.rank = 0; pc.Nl = n; pc.inv_nl = 1.0f / (float)n;
  pc.x[0] = f->fc;
  const int gx = (int)std::min<long long>(nblk(n, 256), 148 * 8);
  dim3 gmean((unsigned)std::min<long long>(nblk(n, 256), 1184), (unsigned)cols), ghist((unsigned)gx, (unsigned)cols);
  k_mean_state<real><<<gmean, 256, 0, f->stream>>>(pc, nullptr, mean_dev, cols, n, f->Ns);
  const int passes = (int)sizeof(real) == 4 ? 4 : 8;
  for (int pass = 0; pass < passes; ++pass) {
    k_select_hist<real><<<ghist, 256, 0, f->stream>>>(pc, nullptr, a, cols, n, f->Ns, pass, sel_dev, hist_dev);
    k_select_pick<<<2 * cols, 32, 0, f->stream>>>(sel_dev, hist_dev);
  }
  k_select_finish<real><<<1, 128, 0, f->stream>>>(sel_dev, out_dev, 2 * cols);
  f->launches += 2 + 2 * passes;
  CU(cudaGetLastError());
  host.resize((size_t)cols + 2 * cols);
  CU(cudaMemcpyAsync(host.data(), mean_dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return CSSM_OK;
}
}  // namespace

extern "C" {

int cssm_version(void) { return CSSM_VERSION; }
const char* cssm_last_error(void) { return g_err.c_str(); }

int cssm_device_count(int* n_out) {
  if (!n_out) return fail(CSSM_ERR_INVALID, "null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *n_out = 0;
    return fail(CSSM_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *n_out = n;
  return CSSM_OK;
}

int cssm_filter_create(const cssm_model_desc_t* model, int64_t n_particles, int resample_kind, int dtype, int device,
                       uint64_t seed, uint64_t stream_id, cssm_filter_t** out) {
  return create_impl(model, n_particles, resample_kind, dtype, device, seed, stream_id, 0, 1, out);
}

int cssm_filter_create_sharded(const cssm_model_desc_t* model, int64_t n_local, int resample_kind, int dtype, int device,
                               uint64_t seed, uint64_t stream_id, int rank, int world, cssm_filter_t** out) {
  return create_impl(model, n_local, resample_kind, dtype, device, seed, stream_id, rank, world, out);
}

int cssm_filter_shard_export(cssm_filter_t* f, void* blob_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!blob_out) return fail(CSSM_ERR_INVALID, "null blob");
  ShardBlob b;
  std::memset(&b, 0, sizeof(b));
  b.magic = BLOB_MAGIC;
  b.rank = f->rank; b.world = f->world; b.device = f->device; b.dtype = f->dtype; b.nt = f->nt; b.d = f->d;
  b.N = f->N; b.Ns = f->Ns; b.pid = (int64_t)getpid();
  void* ptrs[7] = {f->x[0], f->x[1], f->anc, f->logw_base, f->tb.tile_sum, f->tb.tile_maxw, f->xch};
  for (int i = 0; i < 7; ++i) {
    b.ptr[i] = ptrs[i];
    if (f->world > 1) {
      cudaError_t e = cudaIpcGetMemHandle(&b.h[i], ptrs[i]);
      if (e != cudaSuccess) {
        // in-process groups never open the handles; a multi-process job will fail at connect
        cudaGetLastError();
        std::memset(&b.h[i], 0, sizeof(b.h[i]));
      }
    }
  }
  std::memset(blob_out, 0, CSSM_SHARD_BLOB_BYTES);
  std::memcpy(blob_out, &b, sizeof(b));
  return CSSM_OK;
}

int cssm_filter_shard_connect(cssm_filter_t* f, const void* blobs, int world) {
  int rc = enter(f);
  if (rc) return rc;
  if (!blobs || world != f->world) return fail(CSSM_ERR_INVALID, "shard_connect: world size differs from the one the filter was created with");
  if (f->connected) return fail(CSSM_ERR_STATE, "shard_connect: already connected");
  const int64_t mypid = (int64_t)getpid();
  for (int q = 0; q < world; ++q) {
    ShardBlob b;
    std::memcpy(&b, (const char*)blobs + (size_t)q * CSSM_SHARD_BLOB_BYTES, sizeof(b));
    if (b.magic != BLOB_MAGIC || b.rank != q || b.world != world) return fail(CSSM_ERR_INVALID, "shard_connect: malformed blob (expected one blob per rank, in rank order)");
    if (b.N != f->N || b.Ns != f->Ns || b.dtype != f->dtype || b.nt != f->nt || b.d != f->d)
      return fail(CSSM_ERR_INVALID, "shard_connect: ranks disagree on particle count, dtype or model shape");
    if (q == f->rank) continue;
    void* p[7];
    if (b.pid == mypid) {  // same process (one host thread driving several shards): plain pointers
      if (b.device == f->device) f->shared_device = true;
      if (b.device != f->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, f->device, b.device));
        if (!can) return fail(CSSM_ERR_COMM, "shard_connect: no peer access between the devices of two ranks");
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(CSSM_ERR_COMM, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      for (int i = 0; i < 7; ++i) p[i] = b.ptr[i];
    } else {  // another process on this node: CUDA IPC mapping of the peer's allocations (NVLink P2P)
      for (int i = 0; i < 7; ++i) {
        cudaError_t e = cudaIpcOpenMemHandle(&p[i], b.h[i], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(CSSM_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        f->ipc_opened.push_back(p[i]);
      }
    }
    cssm_filter::PeerPtrs& pp = f->peer[q];
    pp.x[0] = p[0]; pp.x[1] = p[1]; pp.anc = (int32_t*)p[2]; pp.logw = p[3];
    pp.tile_sum = (u128*)p[4]; pp.tile_maxw = (double*)p[5]; pp.xch = (XchSlot*)p[6];
  }
  f->connected = true;
  return CSSM_OK;
}

int cssm_filter_shard_info(const cssm_filter_t* f, int32_t* rank_out, int32_t* world_out, int64_t* slot0_out) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  if (rank_out) *rank_out = f->rank;
  if (world_out) *world_out = f->world;
  if (slot0_out) *slot0_out = (int64_t)f->rank * f->N;
  return CSSM_OK;
}

int cssm_filter_set_params(cssm_filter_t* f, const cssm_model_desc_t* model) {
  int rc = enter(f);
  if (rc) return rc;
  HostModel hm;
  rc = copy_model(model, hm);
  if (rc) return rc;
  if (hm.d != f->d || hm.leaves.size() != f->model.leaves.size() || hm.obs_kind != f->model.obs_kind || hm.obs_df != f->model.obs_df)
    return fail(CSSM_ERR_INVALID, "set_params: model shape differs from the one the filter was created with");
  for (size_t l = 0; l < hm.leaves.size(); ++l)
    if (hm.leaves[l].dim != f->model.leaves[l].dim || hm.leaves[l].sde_kind != f->model.leaves[l].sde_kind ||
        hm.leaves[l].f_kind != f->model.leaves[l].f_kind)
      return fail(CSSM_ERR_INVALID, "set_params: leaf shape differs");
  f->model = hm;
  // the constants of a loaded series depend on the parameters: rebuild them
  if (!f->series.empty()) {
    std::vector<double> t, y;
    std::vector<uint8_t> ho;
    for (const SeriesStep& s : f->series) { t.push_back(s.t); y.push_back(s.h.y); ho.push_back((uint8_t)s.h.has_obs); }
    return cssm_filter_load_series(f, t.data(), y.data(), ho.data(), (int64_t)t.size());
  }
  return CSSM_OK;
}

int cssm_filter_reseed(cssm_filter_t* f, uint64_t seed, uint64_t stream_id) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  f->seed = seed; f->stream_id = stream_id; f->epoch = 0;
  return CSSM_OK;
}

int cssm_filter_set_stream(cssm_filter_t* f, void* cuda_stream) {
  int rc = enter(f);
  if (rc) return rc;
  CU(cudaStreamSynchronize(f->stream));
  f->stream = cuda_stream ? (cudaStream_t)cuda_stream : f->own_stream;
  return CSSM_OK;
}

int cssm_filter_destroy(cssm_filter_t* f) {
  if (!f) return CSSM_OK;
  cudaSetDevice(f->device);
  if (f->own_stream) cudaStreamSynchronize(f->own_stream);
  for (void* p : f->ipc_opened) cudaIpcCloseMemHandle(p);
  void* ptrs[] = {f->x[0], f->x[1], f->logw_base, f->anc, f->sc, f->xch, f->tb.tile_sum, f->tb.tile_maxw, f->tb.super_sum, f->tb.super_q,
                  f->tb.super_ticket, f->ubuf, f->cdf, f->scratch, f->ctab, f->ll_steps, f->ess_steps, f->states, f->tile_q, f->series_ctl,
                  f->recs, f->fc, f->px, f->panc, f->pres_dev, f->anc64, f->logw2, f->ser_tile_sum, f->ser_tile_q, f->ser_tile_maxw};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  if (f->pin_ev) cudaEventDestroy(f->pin_ev);
  if (f->pin) cudaFreeHost(f->pin);
  if (f->own_stream) cudaStreamDestroy(f->own_stream);
  delete f;
  return CSSM_OK;
}

int cssm_filter_dim(const cssm_filter_t* f, int32_t* d_out) {
  if (!f || !d_out) return fail(CSSM_ERR_INVALID, "null argument");
  *d_out = f->d;
  return CSSM_OK;
}
int cssm_filter_n_particles(const cssm_filter_t* f, int64_t* n_out) {
  if (!f || !n_out) return fail(CSSM_ERR_INVALID, "null argument");
  *n_out = f->N;
  return CSSM_OK;
}

int cssm_filter_init(cssm_filter_t* f, double t0) {
  int rc = enter(f);
  if (rc) return rc;
  f->launches = 0;
  rc = do_init(f, t0, nullptr, nullptr);
  f->last_launches = f->launches;
  return rc;
}

int cssm_filter_init_state(cssm_filter_t* f, double t0, const double* x0) {
  int rc = enter(f);
  if (rc) return rc;
  if (!x0) return fail(CSSM_ERR_INVALID, "null initial state");
  f->launches = 0;
  rc = do_init(f, t0, nullptr, x0);
  f->last_launches = f->launches;
  return rc;
}

int cssm_filter_init_injected(cssm_filter_t* f, double t0, const double* z0) {
  int rc = enter(f);
  if (rc) return rc;
  if (!z0) return fail(CSSM_ERR_INVALID, "null noise");
  size_t n = (size_t)f->d * f->N;
  rc = ensure_scratch(f, n);
  if (rc) return rc;
  CU(cudaMemcpyAsync(f->scratch, z0, n * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  f->launches = 0;
  rc = do_init(f, t0, f->scratch, nullptr);
  if (rc) return rc;
  CU(cudaStreamSynchronize(f->stream));
  f->last_launches = f->launches;
  return CSSM_OK;
}

int cssm_filter_step(cssm_filter_t* f, double t, int has_obs, double y, double* ll_out, int32_t* ess_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "stepFilter before initialiseState");
  f->launches = 0;
  StepIO io;
  rc = run_one_step(f, t, has_obs, y, io);
  if (rc) return rc;
  f->last_launches = f->launches;
  return read_ll(f, ll_out, ess_out);
}

int cssm_filter_n_substeps(const cssm_filter_t* f, double dt, int64_t* n_out) {
  if (!f || !n_out) return fail(CSSM_ERR_INVALID, "null argument");
  if (f->model.obs_kind != CSSM_OBS_LGCP) { *n_out = 1; return CSSM_OK; }
  *n_out = (dt == 0) ? 0 : (int64_t)(int)std::ceil(dt / std::pow(10, -f->model.lgcp_precision));
  return CSSM_OK;
}

}  // extern "C"

namespace {

// uploads of one injected step: noise rows [n_sub*d] x N taken from a host matrix whose rows have
// `ld` columns starting at column `col0` (a shard's slice of the global noise), and the uniforms
int upload_injected(cssm_filter* f, int64_t n_sub, const double* z, long long ld, long long col0, const double* u, long long n_u,
                    bool weighted, StepIO& io, size_t& nz) {
  nz = (size_t)std::max<int64_t>(n_sub, 1) * f->d * f->N;
  int rc = ensure_scratch(f, nz + (size_t)f->d * f->N);
  if (rc) return rc;
  if (n_sub > 0)
    CU(cudaMemcpy2DAsync(f->scratch, (size_t)f->N * sizeof(double), z + col0, (size_t)ld * sizeof(double), (size_t)f->N * sizeof(double),
                         (size_t)n_sub * f->d, cudaMemcpyHostToDevice, f->stream));
  io.zinj = f->scratch;
  if (weighted) {
    if (f->resample_kind == CSSM_RESAMPLE_SYSTEMATIC) {
      CU(cudaMemcpyAsync(&f->sc->u_inj, u, sizeof(double), cudaMemcpyHostToDevice, f->stream));
      io.use_u_inj = 1;
    } else {
      if (!f->ubuf) CU(cudaMalloc(&f->ubuf, (size_t)n_u * sizeof(double)));
      CU(cudaMemcpyAsync(f->ubuf, u, (size_t)n_u * sizeof(double), cudaMemcpyHostToDevice, f->stream));
      io.uarr = f->ubuf;
    }
  }
  return CSSM_OK;
}

// read-backs of one injected step into host rows of `ld` columns at column col0
int read_injected(cssm_filter* f, size_t nz, bool weighted, long long ld, long long col0, double* x_prop_out, double* logw_out,
                  double* w1_out, int32_t* anc_out) {
  double* stage = f->scratch + nz;  // d*N doubles
  const int g = nblk(f->N, 256);
  if (x_prop_out) {
    if (f->dtype == CSSM_F32) launch_gather<float>(f, nullptr, stage);
    else launch_gather<double>(f, nullptr, stage);
    CU(cudaMemcpy2DAsync(x_prop_out + col0, (size_t)ld * sizeof(double), stage, (size_t)f->N * sizeof(double), (size_t)f->N * sizeof(double),
                         (size_t)f->d, cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
  }
  if (weighted) {
    if (logw_out) {
      if (f->dtype == CSSM_F32) k_to_double<float><<<g, 256, 0, f->stream>>>((const float*)f->logw, stage, f->N);
      else k_to_double<double><<<g, 256, 0, f->stream>>>((const double*)f->logw, stage, f->N);
      CU(cudaMemcpyAsync(logw_out + col0, stage, (size_t)f->N * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
      CU(cudaStreamSynchronize(f->stream));
    }
    if (w1_out) {
      if (f->dtype == CSSM_F32) k_w1_out<float><<<g, 256, 0, f->stream>>>((const float*)f->logw, f->sc, stage, f->N);
      else k_w1_out<double><<<g, 256, 0, f->stream>>>((const double*)f->logw, f->sc, stage, f->N);
      CU(cudaMemcpyAsync(w1_out + col0, stage, (size_t)f->N * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
      CU(cudaStreamSynchronize(f->stream));
    }
    if (anc_out) {
      CU(cudaMemcpyAsync(anc_out + col0, f->anc, (size_t)f->N * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
      CU(cudaStreamSynchronize(f->stream));
    }
  }
  return CSSM_OK;
}

int check_group(cssm_filter_t* const* sh, int R) {
  if (!sh || R < 1 || R > MAXR) return fail(CSSM_ERR_INVALID, "group: bad shard array");
  for (int r = 0; r < R; ++r) {
    if (!sh[r]) return fail(CSSM_ERR_INVALID, "group: null shard");
    if (sh[r]->world != R || sh[r]->rank != r) return fail(CSSM_ERR_INVALID, "group: shards must be passed in rank order, one per rank");
    if (R > 1 && !sh[r]->connected) return fail(CSSM_ERR_STATE, "group: shards are not connected");
  }
  // shards that share a device must share a stream: the lock-step launch order below is what
  // guarantees that every wait inside a kernel is already satisfied when the kernel runs
  for (int r = 1; r < R; ++r)
    for (int q = 0; q < r; ++q)
      if (sh[r]->device == sh[q]->device) { sh[r]->stream = sh[q]->stream; break; }
  return CSSM_OK;
}

#define EACH_SHARD(body)                      \
  for (int r = 0; r < R; ++r) {               \
    cssm_filter* f = sh[r];                   \
    int rc_ = enter(f);                       \
    if (rc_) return rc_;                      \
    body                                      \
  }

// one stepFilter of the whole group in lock-step: K1 on every shard, then K2, then K3
int group_step(cssm_filter_t* const* sh, int R, const StepHost* hs, const long long* n_subs, const void* const* ctabs, double delta,
               const StepIO* ios) {
  StepCtx cx[MAXR];
  EACH_SHARD({
    int rc = (f->dtype == CSSM_F32) ? step_phase1<float>(f, hs[r], n_subs[r], ctabs[r], delta, ios[r], cx[r])
                                    : step_phase1<double>(f, hs[r], n_subs[r], ctabs[r], delta, ios[r], cx[r]);
    if (rc) return rc;
  })
  EACH_SHARD({
    int rc = (f->dtype == CSSM_F32) ? step_phase2<float>(f, cx[r]) : step_phase2<double>(f, cx[r]);
    if (rc) return rc;
  })
  EACH_SHARD({
    int rc = (f->dtype == CSSM_F32) ? step_phase3<float>(f, ios[r], cx[r]) : step_phase3<double>(f, ios[r], cx[r]);
    if (rc) return rc;
  })
  return CSSM_OK;
}

}  // namespace

extern "C" {

int cssm_filter_step_injected(cssm_filter_t* f, double t, int has_obs, double y, const double* z, const double* u,
                              double* x_prop_out, double* logw_out, double* w1_out, int32_t* anc_out, double* ll_out,
                              int32_t* ess_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "stepFilter before initialiseState");
  if (f->world > 1) return fail(CSSM_ERR_UNSUPPORTED, "step_injected on one shard: use cssm_group_step_injected");
  int64_t n_sub = 1;
  cssm_filter_n_substeps(f, t - f->t_cur, &n_sub);
  const bool lgcp = f->model.obs_kind == CSSM_OBS_LGCP;
  const bool weighted = lgcp || has_obs;
  if (n_sub > 0 && !z) return fail(CSSM_ERR_INVALID, "null noise");
  if (weighted && !u) return fail(CSSM_ERR_INVALID, "null uniforms");
  StepIO io;
  size_t nz;
  rc = upload_injected(f, n_sub, z, f->N, 0, u, f->N, weighted, io, nz);
  if (rc) return rc;
  f->launches = 0;
  rc = run_one_step(f, t, has_obs, y, io);
  if (rc) return rc;
  rc = read_injected(f, nz, weighted, f->N, 0, x_prop_out, logw_out, w1_out, anc_out);
  if (rc) return rc;
  f->last_launches = f->launches;
  return read_ll(f, ll_out, ess_out);
}

/* ---- in-process groups: R shards of one filter driven in lock-step by one host thread ---------- */

int cssm_group_init_injected(cssm_filter_t* const* sh, int R, double t0, const double* z0) {
  int rc = check_group(sh, R);
  if (rc) return rc;
  if (!z0) return fail(CSSM_ERR_INVALID, "null noise");
  const long long Ng = (long long)R * sh[0]->N;
  EACH_SHARD({
    size_t n = (size_t)f->d * f->N;
    int rc2 = ensure_scratch(f, n);
    if (rc2) return rc2;
    CU(cudaMemcpy2DAsync(f->scratch, (size_t)f->N * sizeof(double), z0 + (long long)r * f->N, (size_t)Ng * sizeof(double),
                         (size_t)f->N * sizeof(double), (size_t)f->d, cudaMemcpyHostToDevice, f->stream));
    f->launches = 0;
    rc2 = do_init(f, t0, f->scratch, nullptr);
    if (rc2) return rc2;
  })
  EACH_SHARD({ CU(cudaStreamSynchronize(f->stream)); })
  return CSSM_OK;
}

int cssm_group_init(cssm_filter_t* const* sh, int R, double t0) {
  int rc = check_group(sh, R);
  if (rc) return rc;
  EACH_SHARD({
    f->launches = 0;
    int rc2 = do_init(f, t0, nullptr, nullptr);
    if (rc2) return rc2;
  })
  return CSSM_OK;
}

int cssm_group_step_injected(cssm_filter_t* const* sh, int R, double t, int has_obs, double y, const double* z, const double* u,
                             double* x_prop_out, double* logw_out, double* w1_out, int32_t* anc_out, double* ll_out,
                             int32_t* ess_out) {
  int rc = check_group(sh, R);
  if (rc) return rc;
  const long long Ng = (long long)R * sh[0]->N;
  StepHost hs[MAXR];
  long long n_subs[MAXR];
  const void* ctabs[MAXR];
  StepIO ios[MAXR];
  size_t nzs[MAXR];
  double delta = 0.0;
  const bool lgcp = sh[0]->model.obs_kind == CSSM_OBS_LGCP;
  const bool weighted = lgcp || has_obs;
  if (weighted && !u) return fail(CSSM_ERR_INVALID, "null uniforms");
  EACH_SHARD({
    if (!f->initialised) return fail(CSSM_ERR_STATE, "stepFilter before initialiseState");
    int rc2 = prepare_one_step(f, t, has_obs, y, hs[r], n_subs[r], delta, ctabs[r]);
    if (rc2) return rc2;
    if (n_subs[r] > 0 && !z) return fail(CSSM_ERR_INVALID, "null noise");
    rc2 = upload_injected(f, n_subs[r], z, Ng, (long long)r * f->N, u, Ng, weighted, ios[r], nzs[r]);
    if (rc2) return rc2;
    f->launches = 0;
  })
  rc = group_step(sh, R, hs, n_subs, ctabs, delta, ios);
  if (rc) return rc;
  EACH_SHARD({
    f->t_cur = t;
    int rc2 = read_injected(f, nzs[r], weighted, Ng, (long long)r * f->N, x_prop_out, logw_out, w1_out, anc_out);
    if (rc2) return rc2;
    f->last_launches = f->launches;
  })
  double ll = 0.0;
  int32_t ess = 0;
  EACH_SHARD({
    double l;
    int32_t e;
    int rc2 = read_ll(f, &l, &e);
    if (rc2) return rc2;
    if (r == 0) { ll = l; ess = e; }
    else if (std::memcmp(&l, &ll, sizeof(double)) != 0 || e != ess) return fail(CSSM_ERR_COMM, "group: ranks disagree on the log-likelihood");
  })
  if (ll_out) *ll_out = ll;
  if (ess_out) *ess_out = ess;
  return CSSM_OK;
}

int cssm_group_get_particles(cssm_filter_t* const* sh, int R, double* x_out) {
  int rc = check_group(sh, R);
  if (rc) return rc;
  if (!x_out) return fail(CSSM_ERR_INVALID, "null output");
  const long long Ng = (long long)R * sh[0]->N;
  EACH_SHARD({
    if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
    size_t n = (size_t)f->d * f->N;
    int rc2 = ensure_scratch(f, n);
    if (rc2) return rc2;
    const int32_t* anc = f->anc_valid ? f->anc : nullptr;
    if (f->dtype == CSSM_F32) launch_gather<float>(f, anc, f->scratch);
    else launch_gather<double>(f, anc, f->scratch);
    CU(cudaMemcpy2DAsync(x_out + (long long)r * f->N, (size_t)Ng * sizeof(double), f->scratch, (size_t)f->N * sizeof(double),
                         (size_t)f->N * sizeof(double), (size_t)f->d, cudaMemcpyDeviceToHost, f->stream));
  })
  EACH_SHARD({ CU(cudaStreamSynchronize(f->stream)); })
  return CSSM_OK;
}

int cssm_group_ll(cssm_filter_t* const* sh, int R, const double* t, const double* y, const uint8_t* has_obs, int64_t T,
                  double* ll_out, float* ms_out) {
  int rc = check_group(sh, R);
  if (rc) return rc;
  EACH_SHARD({
    int rc2 = cssm_filter_load_series(f, t, y, has_obs, T);
    if (rc2) return rc2;
    rc2 = ensure_steps_cap(f, (size_t)T);
    if (rc2) return rc2;
    f->launches = 0;
    CU(cudaEventRecord(f->ev0, f->stream));
    rc2 = do_init(f, f->t0_series, nullptr, nullptr);
    if (rc2) return rc2;
  })
  StepHost hs[MAXR];
  long long n_subs[MAXR];
  const void* ctabs[MAXR];
  StepIO ios[MAXR];
  for (int64_t s = 0; s < T; ++s) {
    double delta = 0.0;
    for (int r = 0; r < R; ++r) {
      cssm_filter* f = sh[r];
      const SeriesStep& st = f->series[(size_t)s];
      hs[r] = st.h;
      n_subs[r] = st.n_sub;
      ctabs[r] = nullptr;
      if (f->ctab && st.n_sub > 0 && !f->ctab_host.empty())
        ctabs[r] = (const char*)f->ctab + st.ctab_off * ((f->dtype == CSSM_F32) ? 4 : 8);
      delta = (f->model.obs_kind == CSSM_OBS_LGCP) ? std::pow(10, -f->model.lgcp_precision) : 0.0;
      ios[r] = StepIO();
      ios[r].ll_steps = f->ll_steps;
      ios[r].ess_steps = f->ess_steps;
      ios[r].step_slot = (long long)s;
    }
    rc = group_step(sh, R, hs, n_subs, ctabs, delta, ios);
    if (rc) return rc;
    for (int r = 0; r < R; ++r) sh[r]->t_cur = sh[r]->series[(size_t)s].t;
  }
  EACH_SHARD({ CU(cudaEventRecord(f->ev1, f->stream)); })
  double ll = 0.0;
  float ms = 0.f;
  EACH_SHARD({
    double l;
    int rc2 = read_ll(f, &l, nullptr);
    if (rc2) return rc2;
    CU(cudaEventElapsedTime(&f->last_ms, f->ev0, f->ev1));
    f->last_launches = f->launches;
    ms = std::max(ms, f->last_ms);
    if (r == 0) ll = l;
    else if (std::memcmp(&l, &ll, sizeof(double)) != 0) return fail(CSSM_ERR_COMM, "group: ranks disagree on the log-likelihood");
  })
  if (ll_out) *ll_out = ll;
  if (ms_out) *ms_out = ms;
  return CSSM_OK;
}

int cssm_filter_load_series(cssm_filter_t* f, const double* t, const double* y, const uint8_t* has_obs, int64_t T) {
  int rc = enter(f);
  if (rc) return rc;
  if (T <= 0 || !t || !y) return fail(CSSM_ERR_INVALID, "empty series");
  std::vector<SeriesStep> series((size_t)T);
  std::vector<double> ctab_host;
  double t0 = t[0];
  for (int64_t s = 1; s < T; ++s) t0 = std::min(t0, t[s]);  // data.minBy(_.t).t, model/ParticleFilter.scala:138
  double tp = t0;
  const bool lgcp = f->model.obs_kind == CSSM_OBS_LGCP;
  const bool seas = lgcp && has_seasonal(f->model);
  for (int64_t s = 0; s < T; ++s) {
    SeriesStep& st = series[(size_t)s];
    st.t = t[s];
    st.dt = t[s] - tp;
    double delta;
    rc = step_consts(f, tp, t[s], has_obs ? (int)has_obs[s] : 1, y[s], st.h, st.n_sub, delta);
    if (rc) return rc;
    st.ctab_off = ctab_host.size();
    if (seas && st.n_sub > 0) lgcp_ctab_host(f->model, t[s], st.n_sub, delta, ctab_host);
    tp = t[s];
  }
  rc = upload_ctab(f, ctab_host);
  if (rc) return rc;
  f->series.swap(series);
  f->ctab_host.swap(ctab_host);
  f->t0_series = t0;
  f->recs_valid = false;
  return CSSM_OK;
}

int cssm_filter_ll_resident(cssm_filter_t* f, double* ll_out, double* ll_steps_out, int32_t* ess_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (f->series.empty()) return fail(CSSM_ERR_STATE, "no series loaded");
  rc = run_series(f, false);
  if (rc) return rc;
  double ll;
  int32_t ess;
  rc = read_ll(f, &ll, &ess);
  if (rc) return rc;
  CU(cudaEventElapsedTime(&f->last_ms, f->ev0, f->ev1));
  f->last_launches = f->launches;
  if (ll_out) *ll_out = ll;
  const size_t T = f->series.size();
  if (ll_steps_out || ess_out) {
    std::vector<double> lls(T);
    std::vector<int> esss(T);
    CU(cudaMemcpy(lls.data(), f->ll_steps, T * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(esss.data(), f->ess_steps, T * sizeof(int), cudaMemcpyDeviceToHost));
    double pl = 0.0;
    int pe = (int)std::min<long long>(f->N * (long long)f->world, 2147483647LL);
    for (size_t s = 0; s < T; ++s) {
      if (f->series[s].h.has_obs) { pl = lls[s]; pe = esss[s]; }  // unobserved step: ll, ess unchanged (:121)
      if (ll_steps_out) ll_steps_out[s] = pl;
      if (ess_out) ess_out[s] = pe;
    }
  }
  return CSSM_OK;
}

int cssm_filter_series_len(const cssm_filter_t* f, int64_t* T_out) {
  if (!f || !T_out) return fail(CSSM_ERR_INVALID, "null argument");
  *T_out = (int64_t)f->series.size();
  return CSSM_OK;
}

int cssm_filter_ll(cssm_filter_t* f, const double* t, const double* y, const uint8_t* has_obs, int64_t T, double* ll_out) {
  int rc = cssm_filter_load_series(f, t, y, has_obs, T);
  if (rc) return rc;
  return cssm_filter_ll_resident(f, ll_out, nullptr, nullptr);
}

int cssm_filter_run(cssm_filter_t* f, const double* t, const double* y, const uint8_t* has_obs, int64_t T, double* ll_out,
                    double* states_out) {
  int rc = cssm_filter_load_series(f, t, y, has_obs, T);
  if (rc) return rc;
  rc = run_series(f, states_out != nullptr);
  if (rc) return rc;
  double ll;
  rc = read_ll(f, &ll, nullptr);
  if (rc) return rc;
  CU(cudaEventElapsedTime(&f->last_ms, f->ev0, f->ev1));
  f->last_launches = f->launches;
  if (ll_out) *ll_out = ll;
  if (states_out) CU(cudaMemcpy(states_out, f->states, (size_t)(T + 1) * f->d * sizeof(double), cudaMemcpyDeviceToHost));
  return CSSM_OK;
}

int cssm_filter_last_elapsed_ms(const cssm_filter_t* f, float* ms_out) {
  if (!f || !ms_out) return fail(CSSM_ERR_INVALID, "null argument");
  *ms_out = f->last_ms;
  return CSSM_OK;
}
int cssm_filter_last_launches(const cssm_filter_t* f, int64_t* n_out) {
  if (!f || !n_out) return fail(CSSM_ERR_INVALID, "null argument");
  *n_out = f->last_launches;
  return CSSM_OK;
}

int cssm_filter_series_mode(cssm_filter_t* f, int mode) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  if (mode != CSSM_SERIES_AUTO && mode != CSSM_SERIES_THREE_LAUNCH && mode != CSSM_SERIES_SINGLE_LAUNCH)
    return fail(CSSM_ERR_INVALID, "unknown series mode");
  f->series_mode = mode;
  return CSSM_OK;
}

int cssm_filter_profile(cssm_filter_t* f, int stride) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  f->prof_stride = stride > 0 ? stride : 0;
  f->prof_step = 0;
  for (int c = 0; c < 8; ++c) { f->prof_ms[c] = 0; f->prof_n[c] = 0; }
  return CSSM_OK;
}
int cssm_filter_profile_read(cssm_filter_t* f, double* ms_sum_out, int64_t* count_out) {
  if (!f || !ms_sum_out || !count_out) return fail(CSSM_ERR_INVALID, "null argument");
  for (int c = 0; c < 8; ++c) { ms_sum_out[c] = f->prof_ms[c]; count_out[c] = f->prof_n[c]; }
  return CSSM_OK;
}

int cssm_filter_get_particles(cssm_filter_t* f, double* x_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
  if (!x_out) return fail(CSSM_ERR_INVALID, "null output");
  size_t n = (size_t)f->d * f->N;
  rc = ensure_scratch(f, n);
  if (rc) return rc;
  const int32_t* anc = f->anc_valid ? f->anc : nullptr;
  if (f->dtype == CSSM_F32) launch_gather<float>(f, anc, f->scratch);
  else launch_gather<double>(f, anc, f->scratch);
  CU(cudaMemcpyAsync(x_out, f->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return CSSM_OK;
}

int cssm_filter_sample_one(cssm_filter_t* f, double* x_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
  if (!x_out) return fail(CSSM_ERR_INVALID, "null output");
  rc = ensure_scratch(f, (size_t)f->d);
  if (rc) return rc;
  uint32_t tag = 0x40000000u + f->step_ctr++;
  if (f->dtype == CSSM_F32) launch_sample_one<float>(f, f->scratch, tag);
  else launch_sample_one<double>(f, f->scratch, tag);
  CU(cudaMemcpyAsync(x_out, f->scratch, (size_t)f->d * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return CSSM_OK;
}

int cssm_filter_get_ll(cssm_filter_t* f, double* ll_out, int32_t* ess_out) {
  int rc = enter(f);
  if (rc) return rc;
  return read_ll(f, ll_out, ess_out);
}

int cssm_filter_mean_state(cssm_filter_t* f, double* mean_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
  if (!mean_out) return fail(CSSM_ERR_INVALID, "null output");
  rc = ensure_scratch(f, (size_t)f->d);
  if (rc) return rc;
  CU(cudaMemsetAsync(f->scratch, 0, (size_t)f->d * sizeof(double), f->stream));
  dim3 grid((unsigned)std::min<long long>(nblk(f->N, 256), 1184), (unsigned)f->d);
  const int32_t* anc = f->anc_valid ? f->anc : nullptr;
  const Peers pr = make_peers(f, f->cur);
  if (f->dtype == CSSM_F32) k_mean_state<float><<<grid, 256, 0, f->stream>>>(pr, anc, f->scratch, f->d, f->N, f->Ns);
  else k_mean_state<double><<<grid, 256, 0, f->stream>>>(pr, anc, f->scratch, f->d, f->N, f->Ns);
  CU(cudaMemcpyAsync(mean_out, f->scratch, (size_t)f->d * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return CSSM_OK;
}

int cssm_filter_intervals(cssm_filter_t* f, double t, double interval, double* state_mean, double* state_lower,
                          double* state_upper, double* gamma_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
  if (!state_mean || !state_lower || !state_upper || !gamma_out) return fail(CSSM_ERR_INVALID, "null output");
  if (f->world > 1) return fail(CSSM_ERR_UNSUPPORTED, "intervals of a sharded cloud are not implemented");
  const long long n = f->N;
  // getCredibleInterval (model/ParticleFilter.scala:498-503): index = floor(interval * n), sorted(n - index - 1), sorted(index - 1)
  const long long idx_s = (long long)std::floor(interval * (double)n);
  const long long lo_s = n - idx_s - 1, hi_s = idx_s - 1;
  // getOrderStatistic (:455-460): index = floor(n * interval), ordered(n - index), ordered(index)
  const long long idx_e = (long long)std::floor((double)n * interval);
  long long lo_e = n - idx_e, hi_e = idx_e;
  if (idx_e >= n || idx_e < 1) return fail(CSSM_ERR_INVALID, "intervals: the order-statistic index is outside the cloud (the reference throws IndexOutOfBounds here)");
  // decreasing link (BetaModel: exp(-x), model/Model.scala:345): ascending eta is descending gamma, the ranks mirror
  if (f->model.obs_kind == CSSM_OBS_BETA) { lo_e = n - 1 - idx_e; hi_e = idx_e - 1; }
  if (lo_s < 0 || lo_s >= n || hi_s < 0 || hi_s >= n || lo_e < 0 || lo_e >= n || hi_e < 0 || hi_e >= n)
    return fail(CSSM_ERR_INVALID, "intervals: the order-statistic index is outside the cloud (the reference throws IndexOutOfBounds here)");
  const int d = f->d, cols = d + 1;
  // scratch: mean[d] | out[2*cols] | SelState[2*cols] | hist[cols*512 u32]
  const size_t n_dbl = (size_t)d + 2 * cols + 2 * (2 * cols) + (size_t)cols * 256 + 8;
  rc = ensure_scratch(f, n_dbl);
  if (rc) return rc;
  double* mean_dev = f->scratch;
  double* out_dev = mean_dev + d;
  SelState* sel_dev = reinterpret_cast<SelState*>(out_dev + 2 * cols);
  unsigned* hist_dev = reinterpret_cast<unsigned*>(sel_dev + 2 * cols);
  std::vector<SelState> sel((size_t)2 * cols);
  for (int c = 0; c < cols; ++c) {
    sel[2 * c] = SelState{0ull, c < d ? lo_s : lo_e};
    sel[2 * c + 1] = SelState{0ull, c < d ? hi_s : hi_e};
  }
  CU(cudaMemsetAsync(mean_dev, 0, (size_t)d * sizeof(double), f->stream));
  CU(cudaMemsetAsync(hist_dev, 0, (size_t)cols * 512 * sizeof(unsigned), f->stream));
  CU(cudaMemcpyAsync(sel_dev, sel.data(), sel.size() * sizeof(SelState), cudaMemcpyHostToDevice, f->stream));
  const int32_t* anc = f->anc_valid ? f->anc : nullptr;
  const Peers pr = make_peers(f, f->cur);
  StepHost h;
  std::memset(&h, 0, sizeof(h));
  f_coeffs(f->model, t, h.C);
  const int gx = (int)std::min<long long>(nblk(n, 256), 148 * 8);
  dim3 gmean((unsigned)std::min<long long>(nblk(n, 256), 1184), (unsigned)d), ghist((unsigned)gx, (unsigned)cols);
  if (f->dtype == CSSM_F32) {
    StepArgs<float> a;
    to_args<float>(f->model, h, a);
    k_mean_state<float><<<gmean, 256, 0, f->stream>>>(pr, anc, mean_dev, d, n, f->Ns);
    for (int pass = 0; pass < 4; ++pass) {
      k_select_hist<float><<<ghist, 256, 0, f->stream>>>(pr, anc, a, d, n, f->Ns, pass, sel_dev, hist_dev);
      k_select_pick<<<2 * cols, 32, 0, f->stream>>>(sel_dev, hist_dev);
    }
    k_select_finish<float><<<1, 128, 0, f->stream>>>(sel_dev, out_dev, 2 * cols);
  } else {
    StepArgs<double> a;
    to_args<double>(f->model, h, a);
    k_mean_state<double><<<gmean, 256, 0, f->stream>>>(pr, anc, mean_dev, d, n, f->Ns);
    for (int pass = 0; pass < 8; ++pass) {
      k_select_hist<double><<<ghist, 256, 0, f->stream>>>(pr, anc, a, d, n, f->Ns, pass, sel_dev, hist_dev);
      k_select_pick<<<2 * cols, 32, 0, f->stream>>>(sel_dev, hist_dev);
    }
    k_select_finish<double><<<1, 128, 0, f->stream>>>(sel_dev, out_dev, 2 * cols);
  }
  CU(cudaGetLastError());
  std::vector<double> host((size_t)d + 2 * cols);
  CU(cudaMemcpyAsync(host.data(), mean_dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  for (int k = 0; k < d; ++k) {
    state_mean[k] = host[k];
    state_lower[k] = host[d + 2 * k];
    state_upper[k] = host[d + 2 * k + 1];
  }
  gamma_out[0] = host[d + 2 * d];
  gamma_out[1] = host[d + 2 * d + 1];
  return CSSM_OK;
}

int cssm_filter_forecast(cssm_filter_t* f, double t, double interval, int chain, double* state_mean, double* state_lower,
                         double* state_upper, double* eta_out, double* obs_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->initialised) return fail(CSSM_ERR_STATE, "no particles yet");
  if (f->world > 1) return fail(CSSM_ERR_UNSUPPORTED, "forecast of a sharded cloud is not implemented");
  if (f->model.obs_kind == CSSM_OBS_LGCP) return fail(CSSM_ERR_UNSUPPORTED, "LGCP has no observation distribution (`???` in the reference)");
  if (chain && !f->fc_valid) return fail(CSSM_ERR_STATE, "forecast: nothing to continue from");
  const double t_from = chain ? f->fc_t : f->t_cur;
  if (!(t >= t_from)) return fail(CSSM_ERR_INVALID, "forecast: the forecast time lies before the cloud");
  const bool summarise = state_mean || state_lower || state_upper || eta_out || obs_out;
  std::vector<double> host;
  rc = (f->dtype == CSSM_F32) ? forecast_impl<float>(f, t, interval, chain, summarise, host)
                              : forecast_impl<double>(f, t, interval, chain, summarise, host);
  if (rc || !summarise) return rc;
  const int d = f->d, cols = d + 4;
  for (int k = 0; k < d; ++k) {
    if (state_mean) state_mean[k] = host[k];
    if (state_lower) state_lower[k] = host[cols + 2 * k];
    if (state_upper) state_upper[k] = host[cols + 2 * k + 1];
  }
  if (eta_out) { eta_out[0] = host[d + 1]; eta_out[1] = host[cols + 2 * (d + 1)]; eta_out[2] = host[cols + 2 * (d + 1) + 1]; }
  if (obs_out) { obs_out[0] = host[d + 3]; obs_out[1] = host[cols + 2 * (d + 3)]; obs_out[2] = host[cols + 2 * (d + 3) + 1]; }
  return CSSM_OK;
}

int cssm_filter_forecast_cloud(cssm_filter_t* f, double* x_out, double* gamma_out, double* eta_out, double* obs_out,
                               double* obs2_out) {
  int rc = enter(f);
  if (rc) return rc;
  if (!f->fc_valid) return fail(CSSM_ERR_STATE, "no forecast yet");
  const long long n = f->N;
  rc = ensure_scratch(f, (size_t)n);
  if (rc) return rc;
  const int d = f->d;
  for (int c = 0; c < d + 4; ++c) {
    double* dst = c < d ? (x_out ? x_out + (size_t)c * n : nullptr) : c == d ? gamma_out : c == d + 1 ? eta_out : c == d + 2 ? obs_out : obs2_out;
    if (!dst) continue;
    if (f->dtype == CSSM_F32) k_to_double<float><<<nblk(n, 256), 256, 0, f->stream>>>((const float*)f->fc + (size_t)c * f->Ns, f->scratch, n);
    else k_to_double<double><<<nblk(n, 256), 256, 0, f->stream>>>((const double*)f->fc + (size_t)c * f->Ns, f->scratch, n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, f->scratch, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
  }
  return CSSM_OK;
}

int cssm_filter_scan_mode(cssm_filter_t* f, int mode) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  if (mode != CSSM_SCAN_AUTO && mode != CSSM_SCAN_EXACT) return fail(CSSM_ERR_INVALID, "unknown scan mode");
  f->scan_fast = mode == CSSM_SCAN_AUTO ? 1 : 0;
  return CSSM_OK;
}

int cssm_filter_scan_stats(cssm_filter_t* f, int64_t* fast_tiles_out, int64_t* exact_tiles_out) {
  int rc = enter(f);
  if (rc) return rc;
  FilterScalars s;
  CU(cudaMemcpyAsync(&s, f->sc, sizeof(s), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  if (fast_tiles_out) *fast_tiles_out = (int64_t)s.n_fast;
  if (exact_tiles_out) *exact_tiles_out = (int64_t)s.n_exact;
  return CSSM_OK;
}

int cssm_filter_set_tie_rule(cssm_filter_t* f, int rule) {
  if (!f) return fail(CSSM_ERR_INVALID, "null filter handle");
  if (rule != CSSM_TIE_REFERENCE && rule != CSSM_TIE_FIRST) return fail(CSSM_ERR_INVALID, "unknown tie rule");
  f->tie_first = rule == CSSM_TIE_FIRST ? 1 : 0;
  return CSSM_OK;
}

int cssm_filter_paths_enable(cssm_filter_t* f, int64_t max_steps) {
  int rc = enter(f);
  if (rc) return rc;
  if (max_steps < 0) return fail(CSSM_ERR_INVALID, "negative path capacity");
  if (f->world > 1) return fail(CSSM_ERR_UNSUPPORTED, "path storage of a sharded cloud is not implemented");
  CU(cudaStreamSynchronize(f->stream));
  if (f->px) cudaFree(f->px);
  if (f->panc) cudaFree(f->panc);
  if (f->pres_dev) cudaFree(f->pres_dev);
  f->px = nullptr; f->panc = nullptr; f->pres_dev = nullptr;
  f->paths_cap = 0; f->paths_len = -1; f->pres.clear();
  if (max_steps == 0) return CSSM_OK;
  const size_t cloud = (size_t)f->d * f->Ns * (f->dtype == CSSM_F32 ? 4 : 8);
  CU(cudaMalloc(&f->px, (size_t)(max_steps + 1) * cloud));
  CU(cudaMalloc(&f->panc, (size_t)max_steps * f->N * sizeof(int32_t)));
  CU(cudaMalloc(&f->pres_dev, (size_t)max_steps));
  f->paths_cap = max_steps;
  return CSSM_OK;
}

int cssm_filter_paths_len(const cssm_filter_t* f, int64_t* len_out) {
  if (!f || !len_out) return fail(CSSM_ERR_INVALID, "null argument");
  *len_out = f->paths_len;
  return CSSM_OK;
}

int cssm_filter_get_paths(cssm_filter_t* f, const int32_t* idx, int64_t n_idx, double* out) {
  int rc = enter(f);
  if (rc) return rc;
  if (f->paths_cap <= 0 || f->paths_len < 0) return fail(CSSM_ERR_STATE, "no paths recorded (cssm_filter_paths_enable, then initialise)");
  if (!out || n_idx <= 0 || n_idx > f->N) return fail(CSSM_ERR_INVALID, "bad path request");
  if (idx)
    for (int64_t i = 0; i < n_idx; ++i)
      if (idx[i] < 0 || idx[i] >= f->N) return fail(CSSM_ERR_INVALID, "path index outside the cloud");
  const int len = (int)f->paths_len;
  const size_t n_out = (size_t)n_idx * (size_t)(len + 1) * (size_t)f->d;
  const size_t idx_dbl = ((size_t)n_idx * sizeof(int32_t) + 7) / 8;
  rc = ensure_scratch(f, n_out + idx_dbl + 1);
  if (rc) return rc;
  int32_t* idx_dev = nullptr;
  if (idx) {
    idx_dev = reinterpret_cast<int32_t*>(f->scratch + n_out);
    CU(cudaMemcpyAsync(idx_dev, idx, (size_t)n_idx * sizeof(int32_t), cudaMemcpyHostToDevice, f->stream));
  }
  if (len > 0) CU(cudaMemcpyAsync(f->pres_dev, f->pres.data(), (size_t)len, cudaMemcpyHostToDevice, f->stream));
  if (f->dtype == CSSM_F32)
    k_paths<float><<<nblk(n_idx, 256), 256, 0, f->stream>>>((const float*)f->px, f->panc, f->pres_dev, len, f->N, f->Ns, f->d, idx_dev,
                                                             n_idx, f->scratch);
  else
    k_paths<double><<<nblk(n_idx, 256), 256, 0, f->stream>>>((const double*)f->px, f->panc, f->pres_dev, len, f->N, f->Ns, f->d, idx_dev,
                                                              n_idx, f->scratch);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, f->scratch, n_out * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return CSSM_OK;
}

int cssm_resample(int kind, const double* w, int64_t n, const double* u, int64_t n_u, int32_t* ancestors_out, int device) {
  if (kind < 0 || kind > 2) return fail(CSSM_ERR_INVALID, "unknown resample kind");
  if (!w || !u || !ancestors_out || n <= 0 || n > 2147483647LL) return fail(CSSM_ERR_INVALID, "bad resample arguments");
  if (n_u < ((kind == CSSM_RESAMPLE_SYSTEMATIC) ? 1 : n)) return fail(CSSM_ERR_INVALID, "not enough uniforms");
  int ndev = 0;
  int rc = cssm_device_count(&ndev);
  if (rc) return rc;
  if (ndev == 0) return fail(CSSM_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(CSSM_ERR_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  const long long N = n;
  const int items = (N <= (1 << 18)) ? 2 : 8;
  const int tile = TILE_THREADS * items;
  const int nt = nblk(N, tile), ns = nblk(nt, SUPER);
  double *dw = nullptr, *du = nullptr, *dcdf = nullptr;
  int32_t* danc = nullptr;
  FilterScalars* sc = nullptr;
  XchSlot* xch = nullptr;
  SumTables tb = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nt, ns};
  cudaStream_t st = nullptr;
  int status = CSSM_OK;
#define RCU(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess && status == CSSM_OK)                                                    \
      status = fail(CSSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)
  RCU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  RCU(cudaMalloc(&dw, (size_t)(N + tile) * sizeof(double)));
  RCU(cudaMalloc(&du, (size_t)std::max<int64_t>(n_u, 1) * sizeof(double)));
  RCU(cudaMalloc(&danc, (size_t)N * sizeof(int32_t)));
  RCU(cudaMalloc(&sc, sizeof(FilterScalars)));
  RCU(cudaMalloc(&xch, sizeof(XchSlot)));
  RCU(cudaMalloc(&tb.tile_sum, (size_t)(nt + 1) * sizeof(u128)));
  RCU(cudaMalloc(&tb.tile_maxw, (size_t)(nt + 1) * sizeof(double)));
  RCU(cudaMalloc(&tb.super_sum, (size_t)2 * ns * sizeof(u128)));
  RCU(cudaMalloc(&tb.super_q, (size_t)2 * ns * sizeof(u128)));
  RCU(cudaMalloc(&tb.super_ticket, (size_t)ns * sizeof(unsigned long long)));
  if (kind == CSSM_RESAMPLE_MULTINOMIAL) RCU(cudaMalloc(&dcdf, (size_t)N * sizeof(double)));
  if (status == CSSM_OK) {
    FilterScalars z;
    std::memset(&z, 0, sizeof(z));
    z.qb = 96;
    if (kind == CSSM_RESAMPLE_SYSTEMATIC) z.u_inj = u[0];
    RCU(cudaMemsetAsync(dw, 0, (size_t)(N + tile) * sizeof(double), st));
    RCU(cudaMemcpyAsync(dw, w, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st));
    RCU(cudaMemcpyAsync(du, u, (size_t)((kind == CSSM_RESAMPLE_SYSTEMATIC) ? 1 : N) * sizeof(double), cudaMemcpyHostToDevice, st));
    RCU(cudaMemcpyAsync(sc, &z, sizeof(z), cudaMemcpyHostToDevice, st));
    RCU(cudaMemsetAsync(tb.super_sum, 0, (size_t)2 * ns * sizeof(u128), st));
    RCU(cudaMemsetAsync(tb.super_q, 0, (size_t)2 * ns * sizeof(u128), st));
    RCU(cudaMemsetAsync(tb.super_ticket, 0, (size_t)ns * sizeof(unsigned long long), st));
    Peers pr;
    std::memset(&pr, 0, sizeof(pr));
    pr.R = 1; pr.rank = 0; pr.Nl = N; pr.inv_nl = 1.0f / (float)N; pr.anc[0] = danc; pr.xch[0] = xch; pr.tile_sum[0] = tb.tile_sum; pr.tile_maxw[0] = tb.tile_maxw;
    K3Ctl ctl;
    std::memset(&ctl, 0, sizeof(ctl));
    ctl.inv_n = ((N & (N - 1)) == 0) ? 1.0 / (double)N : 0.0;
    ctl.direct = 1; ctl.add_ll = 0; ctl.use_u_inj = 1;
    const bool multi = kind == CSSM_RESAMPLE_MULTINOMIAL;
    const bool strat = kind == CSSM_RESAMPLE_STRATIFIED;
    const double* ua = strat ? du : nullptr;
    double* cdf = multi ? dcdf : nullptr;
    k_max_direct<<<std::min(nblk(N, 256), 1184), 256, 0, st>>>(dw, N, sc);
#define RS_CASE(IT, KD)                                                                   \
  do {                                                                                    \
    k_weight_sums<double, IT><<<nt, TILE_THREADS, 0, st>>>(nullptr, dw, N, sc, 0, 0ull, tb, pr); \
    k_scan_search<double, IT, KD><<<nt, TILE_THREADS, 0, st>>>(nullptr, dw, N, sc, tb, pr, ctl, ua, cdf); \
  } while (0)
    if (items == 8) { if (strat) RS_CASE(8, CSSM_RESAMPLE_STRATIFIED); else RS_CASE(8, CSSM_RESAMPLE_SYSTEMATIC); }
    else { if (strat) RS_CASE(2, CSSM_RESAMPLE_STRATIFIED); else RS_CASE(2, CSSM_RESAMPLE_SYSTEMATIC); }
#undef RS_CASE
    if (multi) k_multinomial_search<<<nblk(N, 256), 256, 0, st>>>(dcdf, N, du, 0u, 0u, 0u, danc, &sc->flags);
    RCU(cudaGetLastError());
    RCU(cudaMemcpyAsync(ancestors_out, danc, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RCU(cudaStreamSynchronize(st));
  }
#undef RCU
  void* ptrs[] = {dw, du, danc, sc, xch, tb.tile_sum, tb.tile_maxw, tb.super_sum, tb.super_q, tb.super_ticket, dcdf};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (st) cudaStreamDestroy(st);
  return status;
}

}  // extern "C"
