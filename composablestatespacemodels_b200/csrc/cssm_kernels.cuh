// cssm_kernels.cuh -- the sm_100a kernels of the particle-filter hot path.
//
// One observed stepFilter (model/ParticleFilter.scala:116-132) is THREE launches:
//
//   k_propagate_weight  K1  gather of the previous resampling (fused into the load, over NVLink
//                           when the ancestor lives on another rank) + stepFunction + f +
//                           dataLikelihood + block max            model/ParticleFilter.scala:118,123-124
//     (k_lgcp_weight    K1' FilterLgcp.calcWeight                 model/ParticleFilter.scala:184-217)
//   k_weight_sums       K2  w1 = exp(logw - max), exact per-tile / per-super-tile / total sums of
//                           w1 and w1^2                           :125, :522-524, :431-434
//   k_scan_search       K3  tile prefix (from the super-tile sums, no separate scan launch),
//                           inclusive CDF of the tile in shared memory, ancestor search with the
//                           TreeMap duplicate-key rule, ll += max + log(mean w1), ESS
//                                                                 model/Resampling.scala:36-86, PF :127-128
//   k_multinomial_search K4' per-draw inverse CDF                 model/Resampling.scala:92-96
//   k_init_particles, k_gather, k_sample_one, k_mean_state        PF :105-108, Resampling :42,:95,:151-154
//
// Layout: structure of arrays, x[k*Ns + i] (Ns = N rounded up to 64) -- every access is unit-stride
// across a warp except the ancestor gather, whose indices are non-decreasing for systematic /
// stratified resampling.  All weight sums are exact 2^-96 fixed point (cssm_common.cuh), so no
// result depends on the tile size, the grid or the number of GPUs a cloud is sharded over.
//
// Sharded filters (R ranks, one GPU each, N_l particles per rank): the same three kernels.  The
// three per-step exchanges -- max log-weight, (sum w, sum w^2), "resampling done" -- are 8..32 byte
// stores into the peers' memory with a release flag, written by the last block of the producing
// kernel and awaited by the first instruction of the consuming kernel (an all-gather without a
// collective launch).  K3 scatters ancestor indices to the rank that owns the offspring slot and
// the next K1 gathers the parent state from the rank that owns the parent: both are plain
// st.global / ld.global on peer-mapped pointers.
#pragma once
#include <type_traits>
#include "cssm_common.cuh"
#include "../../include/cssm.h"

namespace cssm {

// Optimiser sensitivity.  For IDENTICAL source ptxas emits different code for the step kernels depending on what else the
// translation unit contains (round 2: k_propagate_weight<float,7> 82 KB vs 101 KB of code, 0.209 vs 0.191 ms at 2^24
// particles; k_weight_sums 0.042 vs 0.049 ms the other way round; the scan + search and the series kernel likewise --
// reproducible per build and identical on different B200 boxes, profiles/r02_summary.md).  The build is deterministic,
// so the choice can be pinned: CSSM_LAYOUT_PAD adds a never-launched kernel of that many KiB to the module, which is
// enough to move the optimiser from one variant to the other, and scripts/gpu_layout.sh measures the candidates.  The
// default is the one that measured best for this tree.
#ifndef CSSM_LAYOUT_PAD
#define CSSM_LAYOUT_PAD 0  // the build passes -DCSSM_LAYOUT_PAD=<k> (__graft_entry__.LAYOUT_PAD)
#endif
#if CSSM_LAYOUT_PAD > 0
static __global__ void k_layout_pad(float* p) {
  float v = p[0];
#pragma unroll
  for (int i = 0; i < CSSM_LAYOUT_PAD * 64; ++i) v = fmaf(v, 1.0001f, 0.5f);  // 16 bytes of code each
  p[0] = v;
}
#endif

constexpr int MAXD = 32;          // by-value kernel argument budget (5*32*8 B = 1.25 KB)
constexpr int MAXR = CSSM_MAX_RANKS;
constexpr int TILE_THREADS = 256;
constexpr int SUPER = 256;        // tiles per super tile (second level of the prefix)

// per-observation constants, built on the host in fp64 and rounded to the filter dtype.
// transition of coordinate k:  x' = A*(x - M) + M + D + S*z     (exact or Euler-Maruyama),
// evaluated as fma(S, z, fma(A, x, B)) with B = M - A*M + D formed on the host in fp64 (StepArgs.D
// carries B; StepArgs.M is only used by the initial-state kernels)
//   Brownian      exact: A=1 M=0 D=0      S=sqrt(sigma*dt)        EM: D=dt        S=sigma*sqrt(dt)
//   GenBrownian   exact: A=1 M=0 D=mu*dt  S=sqrt(sigma*dt)        EM: D=mu*dt     S=sigma*sqrt(dt)
//   OU            exact: A=exp(-phi dt) M=mu D=0 S=sqrt(sigma^2/(2phi)(1-exp(-2phi dt)))
//                 EM:    A=1-phi*dt     M=mu D=0 S=sigma*sqrt(dt)
// gamma = sum_k C[k]*x'[k]   (C = 1 on the first coordinate of a first-component leaf, the
// cos/sin row of SeasonalModel.buildF on a seasonal leaf, 0 elsewhere)
template <typename real>
struct StepArgs {
  real A[MAXD], M[MAXD], D[MAXD], S[MAXD], C[MAXD];
  real y, k0, k1, k2, k3;  // observation constants, see obs_loglik
  int d, obs_kind, has_obs;
};

// accumulators of one observed step; double-buffered by the parity of the observed-step counter,
// the buffer of the NEXT observed step is zeroed by K3
struct StepAcc {
  unsigned long long gmax_key;  // ordered key of max(logw), atomicMax target (0 = below everything)
  u128 tot;                     // sum fix(w1, qb)          (this rank)
  u128 q;                       // sum fix(w1^2, 96)        (this rank)
  unsigned long long pad;
};
// device-resident scalars of one filter
struct FilterScalars {
  StepAcc acc[2];
  unsigned long long ticket1, ticket2, ticket3;  // monotone "last block" tickets of K1 / K2 / K3
  unsigned long long gate1, gate2, gate3;        // sharded: highest exchange value already seen complete by a block
  double ll, ll_incr, u_inj;
  double gmax, total;  // of the last observed step (read-back of w1, tests)
  int ess, flags, qb, pad;
  unsigned long long n_fast, n_exact;  // tiles of k_scan_search settled by the certified fp64 path / handed to the exact path
  // sharded: the global output slots [out_lo, out_hi) of the last resampling belong to THIS rank's particles, so this rank's
  // own search wrote their ancestors (block 0 of K3 writes the pair, the next K1 reads it)
  long long out_lo, out_hi;
};
enum : int { FLAG_NAN_WEIGHT = 1, FLAG_ZERO_TOTAL = 2, FLAG_CLAMPED = 4, FLAG_COMM_TIMEOUT = 8 };

// what rank q writes into the memory of every rank (slot [q] of the destination's array)
struct XchSlot {
  unsigned long long max_key[2], max_seq[2];                  // after K1
  unsigned long long tot_lo[2], tot_hi[2], q_lo[2], q_hi[2];  // after K2
  unsigned long long sum_seq[2];
  unsigned long long progress;  // number of completed steps (after K3, or after K1 of an unobserved step)
  unsigned long long pad;
};
static_assert(sizeof(XchSlot) == 128, "XchSlot is one 128-byte line");

// the ranks of a sharded filter as one kernel argument (R == 1: the plain single-GPU filter)
struct Peers {
  int R, rank;
  long long Nl;                  // particles per rank (the same on every rank)
  float inv_nl;                  // ~1/Nl (owner_of)
  const void* x[MAXR];           // cloud each rank wrote in the previous step (K1 gathers from it)
  int32_t* anc[MAXR];            // ancestor buffers (K3 scatters into them)
  const void* logw[MAXR];        // log-weights, tile sums, tile maxima: only read when a run of
  const u128* tile_sum[MAXR];    // duplicate keys crosses a rank boundary
  const double* tile_maxw[MAXR];
  XchSlot* xch[MAXR];            // xch[q] = slot array in rank q's memory; this rank writes xch[q][rank]
};

// rank that owns global particle / slot index g (< 2^31): g / Nl without an integer division.  The
// quotient is at most MAXR - 1, so the fp32 estimate is within one of it and one correction settles it.
__device__ __forceinline__ unsigned owner_of(const Peers& pr, unsigned g) {
  unsigned q = __float2uint_rz(__uint2float_rz(g) * pr.inv_nl);
  const int r = (int)(g - q * (unsigned)pr.Nl);
  if (r < 0) --q;
  else if (r >= (int)pr.Nl) ++q;
  return q;
}

// ---------------------------------------------------------------------------------------------
template <typename real> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; static constexpr int PPT = 4; };
template <> struct VecOf<double> { typedef double2 type; static constexpr int PPT = 2; };

template <typename real> __device__ __forceinline__ real r_exp(real x);
template <> __device__ __forceinline__ float r_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double r_exp<double>(double x) { return exp(x); }
template <typename real> __device__ __forceinline__ real r_log(real x);
template <> __device__ __forceinline__ float r_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double r_log<double>(double x) { return log(x); }
template <typename real> __device__ __forceinline__ real r_fma(real a, real b, real c);
template <> __device__ __forceinline__ float r_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double r_fma<double>(double a, double b, double c) { return fma(a, b, c); }
template <typename real> __device__ __forceinline__ real r_add(real a, real b);
template <> __device__ __forceinline__ float r_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double r_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename real> __device__ __forceinline__ real r_sub(real a, real b);
template <> __device__ __forceinline__ float r_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double r_sub<double>(double a, double b) { return __dsub_rn(a, b); }
template <typename real> __device__ __forceinline__ real r_mul(real a, real b);
template <> __device__ __forceinline__ float r_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double r_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename real> __device__ __forceinline__ real r_neg_big();
// Bernoulli "impossible" value: -1e99 in the reference (model/Model.scala:332,334); fp32 cannot
// hold it, the largest finite negative is used instead
template <> __device__ __forceinline__ float r_neg_big<float>() { return -3.4028234663852886e38f; }
template <> __device__ __forceinline__ double r_neg_big<double>() { return -1e99; }

// a7: dataLikelihood of the left-most model.  Host-hoisted constants:
//   POISSON   k0 = k = y.toInt, k1 = lgamma(k+1)                      model/Model.scala:269-273
//   NEGBIN    k0 = k, k1 = size = exp(scale), k2 = lgamma(size+k)-lgamma(k+1)-lgamma(size),
//             k3 = log(size)                                           :186-195
//   NORMAL    k0 = sd = exp(scale), k1 = log(sqrt(2 pi)) + log(sd)     :227-233,:252-258
//   BERNOULLI k0 = (y == 1.0)                                          :318-336
//   STUDENT_T, ZIP, BETA: see the cases                                :144-162, :281-309, :339-353
// Written with explicit round-to-nearest operations only: the compiler never contracts intrinsics
// into FMAs, so every instantiation of every kernel (per-step, series) evaluates the same sequence
// and returns the same bits.
template <typename real>
__device__ __forceinline__ real obs_loglik(const StepArgs<real>& a, real g) {
  switch (a.obs_kind) {
    case CSSM_OBS_POISSON:
      return r_sub<real>(r_fma<real>(a.k0, g, -r_exp<real>(g)), a.k1);  // -mean + k*log(mean) - lgamma(k+1), log(exp(g)) = g
    case CSSM_OBS_NEGBIN: {
      const real L = r_log<real>(r_add<real>(r_exp<real>(g), a.k1));  // log(mu + size)
      return r_fma<real>(a.k0, r_sub<real>(g, L), r_fma<real>(a.k1, r_sub<real>(a.k3, L), a.k2));
    }
    case CSSM_OBS_NORMAL: {
      const real dd = r_sub<real>(a.y, g) / a.k0;
      return r_fma<real>(r_mul<real>((real)-0.5, dd), dd, -a.k1);
    }
    case CSSM_OBS_BERNOULLI: {
      const real p = (g > (real)6) ? (real)1 : (g < (real)-6) ? (real)0 : (real)1 / r_add<real>((real)1, r_exp<real>(-g));
      if (a.k0 != (real)0) return (p == (real)0) ? r_neg_big<real>() : r_log<real>(p);
      return (p == (real)1) ? r_neg_big<real>() : r_log<real>(r_sub<real>((real)1, p));
    }
    case CSSM_OBS_STUDENT_T: {  // k0 = v, k1 = df, k2 = -logNormalizer, k3 = (df+1)/2
      const real xx = r_sub<real>(a.y, g) / a.k0;
      const real L = r_log<real>(r_add<real>((real)1, r_mul<real>(xx, xx) / a.k1));
      return r_fma<real>(-a.k3, L, a.k2) / a.k0;
    }
    case CSSM_OBS_ZIP: {  // k0 = k, k1 = lgamma(k+1), k2 = p, k3 = log(1 + e^v)
      const real lam = r_exp<real>(g);
      if (a.k0 == (real)0) return r_log<real>(r_fma<real>(r_sub<real>((real)1, a.k2), r_exp<real>(-lam), a.k2));
      return r_sub<real>(r_sub<real>(r_fma<real>(a.k0, g, -a.k3), lam), a.k1);
    }
    case CSSM_OBS_BETA:  // k0 = log y, k1 = 0*log(1-y):  (e^-g - 1) log y + 0*log(1-y) + log(e^-g)
      return r_fma<real>(r_sub<real>(r_exp<real>(-g), (real)1), a.k0, r_sub<real>(a.k1, g));
    default: return (real)0;
  }
}

// four N(0,1) per Philox block in fp32, two in fp64
template <typename real> struct Normals;
template <> struct Normals<float> {
  static constexpr int PER_CALL = 4;
  __device__ __forceinline__ static void draw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, float* z) {
    uint4 v = philox4x32(make_uint4(c0, c1, c2, c3), k0, k1);
    box_muller_f(v.x, v.y, z[0], z[1]);
    box_muller_f(v.z, v.w, z[2], z[3]);
  }
};
template <> struct Normals<double> {
  static constexpr int PER_CALL = 2;
  __device__ __forceinline__ static void draw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, double* z) {
    uint4 v = philox4x32(make_uint4(c0, c1, c2, c3), k0, k1);
    box_muller_d(v, z[0], z[1]);
  }
};

// ---------------------------------------------------------------------------------------------
// exchange helpers (sharded filters only; every call site is behind `pr.R > 1`)
// ---------------------------------------------------------------------------------------------
// spin until *p >= want.  Bounded: after 4 s the filter is flagged and every later wait returns
// at once, so a dead peer costs seconds, not a hung GPU.
__device__ __forceinline__ void wait_ge(const unsigned long long* p, unsigned long long want, FilterScalars* sc) {
  if (ld_relaxed_sys(p) >= want) return;
  if (*(volatile int*)&sc->flags & FLAG_COMM_TIMEOUT) return;
  const unsigned long long t0 = global_timer_ns();
  while (ld_relaxed_sys(p) < want) {
    __nanosleep(32);
    if (global_timer_ns() - t0 > 4000000000ull) {
      atomicOr(&sc->flags, FLAG_COMM_TIMEOUT);
      return;
    }
  }
}
// First warp of a block: wait until the flag every peer q keeps in our memory (flag_of(q)) has
// reached `want`.  Lane q polls peer q, so the R waits overlap; and once one block has seen all of
// them the per-filter gate lets every later block through with a single L2 read.  No load of the
// exchanged data (or of peer clouds) is issued by any block before it has passed here, and L1 is
// clean at kernel entry, so the relaxed polling cannot be followed by a stale read.
template <typename FlagOf>
__device__ __forceinline__ void gate_wait(unsigned long long* gate, unsigned long long want, const Peers& pr, FilterScalars* sc,
                                          FlagOf flag_of) {
  if (threadIdx.x < 32) {
    const unsigned long long seen = __shfl_sync(0xffffffffu, ld_gpu(gate), 0);  // warp-uniform decision
    if (seen < want) {
      if ((int)threadIdx.x < pr.R) wait_ge(flag_of((int)threadIdx.x), want, sc);
      __syncwarp();
      if (threadIdx.x == 0) {
        (void)ld_acquire_sys(flag_of(pr.rank));  // orders this block's later reads after the flags
        atomicMax(gate, want);
      }
    }
  }
  __syncthreads();
}
// The first warp of a block publishes to the peers: lane q writes into rank q's memory, so the R
// NVLink round trips of the release stores overlap instead of following one another (a serial loop
// over 8 peers cost ~25 us per exchange, three times per observation).  `last` is lane 0's verdict
// ("this was the last block of the launch"); all 32 lanes call.  Every rank's slot array gets the
// flag, our own included: gate_wait polls all R slots alike.
__device__ __forceinline__ bool warp_is_last(int last_lane0) { return __shfl_sync(0xffffffffu, last_lane0, 0) != 0; }
__device__ __forceinline__ void push_progress_warp(const Peers& pr, unsigned long long gstep_done) {
  __threadfence_system();
  const int q = threadIdx.x & 31;
  if (q < pr.R) st_release_sys(&pr.xch[q][pr.rank].progress, gstep_done);
}
// Ranks that share one device also share one stream (their kernels spin on each other's flags and CUDA does not
// co-schedule streams): there the consumer of rank 0 runs before the producer of rank 1 has published, so a one-warp
// kernel behind each producing kernel publishes instead (k_publish, below the helpers).
// Consumer-side publication (sharded filters; first warp of block 0 of the consuming kernel, before its own gate): the
// producing kernel is complete -- the launches of a sharded filter are serialised by the stream -- so its result is final
// in this rank's memory and one warp hands it to the peers.
//   K1 of step s+1: "every kernel of the steps before is complete" (the peers' searches may have written our ancestors)
//   K2: this rank's max log-weight;   K3: this rank's exact (sum w, sum w^2)
__device__ __forceinline__ void publish_max_warp(const Peers& pr, FilterScalars* sc, int parity, unsigned long long obs_seq) {
  const unsigned long long key = ld_gpu(&sc->acc[parity].gmax_key);
  __threadfence_system();
  const int q = threadIdx.x & 31;
  if (q < pr.R) {
    XchSlot* s = &pr.xch[q][pr.rank];
    st_relaxed_sys(&s->max_key[parity], key);
    st_release_sys(&s->max_seq[parity], obs_seq + 1);
  }
}
__device__ __forceinline__ void publish_sums_warp(const Peers& pr, FilterScalars* sc, int parity, unsigned long long obs_seq) {
  const StepAcc* A = &sc->acc[parity];
  const unsigned long long tl = ld_gpu(&A->tot.lo), th = ld_gpu(&A->tot.hi), ql = ld_gpu(&A->q.lo), qh = ld_gpu(&A->q.hi);
  __threadfence_system();
  const int q = threadIdx.x & 31;
  if (q < pr.R) {
    XchSlot* s = &pr.xch[q][pr.rank];
    st_relaxed_sys(&s->tot_lo[parity], tl);
    st_relaxed_sys(&s->tot_hi[parity], th);
    st_relaxed_sys(&s->q_lo[parity], ql);
    st_relaxed_sys(&s->q_hi[parity], qh);
    st_release_sys(&s->sum_seq[parity], obs_seq + 1);
  }
}

// what: 0 = "steps complete" (value = their number), 1 = this rank's max log-weight, 2 = its exact sums
static __global__ void __launch_bounds__(32) k_publish(const __grid_constant__ Peers pr, FilterScalars* sc, int what, int parity,
                                                       unsigned long long value) {
  if (what == 0) push_progress_warp(pr, value);
  else if (what == 1) publish_max_warp(pr, sc, parity, value);
  else publish_sums_warp(pr, sc, parity, value);
}

// ---------------------------------------------------------------------------------------------
// K0  x0[k][i] = sqrt(c0_k) * z + m0_k        (a.S = sqrt(c0), a.M = m0)
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) k_init_particles(StepArgs<real> a, real* __restrict__ x,
                                                        const double* __restrict__ zinj, long long N,
                                                        long long Ns, unsigned long long slot0, uint32_t key0,
                                                        uint32_t key1, uint32_t epoch) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned long long slot = slot0 + (unsigned long long)i;
  constexpr int PC = Normals<real>::PER_CALL;
  for (int kk = 0; kk < a.d; kk += PC) {
    real z[PC];
    if (zinj == nullptr) Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), epoch, RNG_INIT | (uint32_t)(kk / PC), key0, key1, z);
#pragma unroll
    for (int j = 0; j < PC; ++j) {
      int k = kk + j;
      if (k < a.d) {
        real zz = zinj ? (real)zinj[(long long)k * N + i] : z[j];
        x[(long long)k * Ns + i] = r_fma<real>(a.S[k], zz, a.M[k]);
      }
    }
  }
}

// all particles = one given state (FilterInit, model/ParticleFilter.scala:257-260); a.M = x0
template <typename real>
__global__ void __launch_bounds__(256) k_fill_particles(StepArgs<real> a, real* __restrict__ x, long long N, long long Ns) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int k = 0; k < a.d; ++k) x[(long long)k * Ns + i] = a.M[k];
}

// ---------------------------------------------------------------------------------------------
// tail of K1 / K1': block max -> one atomicMax per block; sharded: the last block of the grid
// publishes the rank's max (observed step) or its progress (unobserved step) to the peers
// ---------------------------------------------------------------------------------------------
struct K1Ctl {
  FilterScalars* sc;
  int parity;                   // observed-step parity
  unsigned long long obs_seq;   // observed steps completed before this one
  unsigned long long gstep;     // steps completed before this one
  int local_ok;                 // sharded: blocks whose ancestors this rank's own search wrote may start without the peers
  int pub_here;                 // sharded: block 0 of this kernel publishes what the kernel before it left (else k_publish did)
};
// (Round 2 measured a tail without block barriers -- every warp folds its max into a shared word with an atomic and the warp
// that counts in last publishes: K1 0.2064 vs 0.2038 ms with Philox4x32-10, 0.1908 vs 0.1908 with 7 rounds.  No gain; this
// simpler form stays.)
template <bool SH>
__device__ __forceinline__ void k1_tail(double mx, bool bad, int has_obs, const Peers& pr, const K1Ctl& ctl) {
  const int RK = SH ? pr.R : 1;
  __shared__ double s_mx[8];
  __shared__ int s_bad;
  FilterScalars* sc = ctl.sc;
  if (has_obs) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
    if (bad) s_bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m2 = s_mx[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m2 = fmax(m2, s_mx[w]);
      atomicMax(&sc->acc[ctl.parity].gmax_key, ord_key(m2));
      if (s_bad) atomicOr(&sc->flags, FLAG_NAN_WEIGHT);
    }
  }
  // Sharded: nothing is published here.  The kernels of a sharded filter run in stream order, so block 0 of the NEXT
  // kernel finds this one complete and publishes its result to the peers (publish_* below): no fence, no ticket atomic
  // and no wait for its answer at the end of every block of this kernel.
  (void)RK;
}

// ---------------------------------------------------------------------------------------------
// K1  fused gather + propagate + f + log-weight + max
//     one thread = PPT consecutive particles (16-byte stores); the loads go through the ancestor
//     index (non-decreasing for systematic/stratified, so a warp still touches few sectors).
//     D > 0: latent dimension known at compile time (everything unrolls, constants become
//     immediate constant-bank operands); D == 0: any d <= MAXD.
// ---------------------------------------------------------------------------------------------
template <typename real, int PPT> struct VecN;
template <> struct VecN<float, 4> { typedef float4 type; typedef int4 itype; };
template <> struct VecN<float, 2> { typedef float2 type; typedef int2 itype; };
template <> struct VecN<double, 2> { typedef double2 type; typedef int2 itype; };
template <> struct VecN<float, 1> { typedef float type; typedef int itype; };
template <> struct VecN<double, 1> { typedef double type; typedef int itype; };

// the PPT consecutive particles i0 .. i0+PPT-1 of one thread: gather, transition, f, log-weight.
// Returns the thread's max log-weight (mx, -inf without an observation) and whether one was NaN.
// Shared by the per-step kernel below and by the single-launch series kernel (cssm_series.cuh), so
// both evaluate a particle with the same instruction sequence.  COH: the cloud and the ancestors
// were written earlier in the SAME launch by other blocks, so they are read with ld.global.cg (L2)
// instead of the non-coherent read-only path.
// the normals of the coordinates kk .. kk+3 of the thread's PPT particles: counter = (global slot, step, chunk)
template <typename real, int PPT>
__device__ __forceinline__ void chunk_noise(int d, int kk, unsigned long long slot0, long long i0, uint32_t step, uint32_t key0,
                                            uint32_t key1, real (&z)[4][PPT]) {
  constexpr int PC = Normals<real>::PER_CALL;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    unsigned long long slot = slot0 + (unsigned long long)(i0 + p);
    real zz[4];
    if (PC == 4) {
      Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)(kk >> 2), key0, key1, zz);
    } else {
      Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)(kk >> 1), key0, key1, zz);
      if (kk + 2 < d)
        Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)((kk >> 1) + 1), key0, key1, zz + 2);
      else
        zz[2] = zz[3] = (real)0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j][p] = zz[j];
  }
}

// sidx != NULL: the PPT source slots, already resolved by the caller (series kernel: polled from tagged ancestor words);
// zpre != NULL (D > 0 only): the step's normals, drawn ahead by chunk_noise into zpre[chunk][4][PPT]
template <typename real, int D, int PPT, bool COH = false, bool FULLBLK = false, bool SH = false>
__device__ __forceinline__ void propagate_particles(const StepArgs<real>& a, const Peers& pr, real* __restrict__ xdst,
                                                    const int32_t* __restrict__ anc, real* __restrict__ logw,
                                                    const double* __restrict__ zinj, long long N, long long Ns,
                                                    unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step,
                                                    long long i0, double& mx, bool& bad, real* lw_out = nullptr,
                                                    const int* sidx = nullptr, const real (*zpre)[4][PPT] = nullptr) {
  typedef typename VecN<real, PPT>::type vec_t;
  typedef typename VecN<real, PPT>::itype ivec_t;
  const int d = (D > 0) ? D : a.d;
  const bool full = FULLBLK || (i0 + PPT <= N);  // FULLBLK: the caller knows that every thread of the block is full
  const real* src[PPT];
  bool valid[PPT];
  {
    const real* xloc = reinterpret_cast<const real*>(pr.x[pr.rank]);
    long long s[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      valid[p] = FULLBLK || (i0 + p < N);
      s[p] = i0 + p;
    }
    if (sidx != nullptr) {
#pragma unroll
      for (int p = 0; p < PPT; ++p) s[p] = valid[p] ? (long long)sidx[p] : i0 + p;
    } else if (anc != nullptr) {
      if (full) {
        const ivec_t v = COH ? __ldcg(reinterpret_cast<const ivec_t*>(anc + i0)) : *reinterpret_cast<const ivec_t*>(anc + i0);
        const int* vp = reinterpret_cast<const int*>(&v);
#pragma unroll
        for (int p = 0; p < PPT; ++p) s[p] = vp[p];
      } else {
#pragma unroll
        for (int p = 0; p < PPT; ++p)
          if (valid[p]) s[p] = COH ? __ldcg(anc + i0 + p) : anc[i0 + p];
      }
    }
    if (SH && pr.R > 1 && anc != nullptr) {
      // ancestors are GLOBAL particle indices.  With exchangeable particles nearly every parent lives on this rank (only
      // the offspring around the rank borders migrate): one range test per particle decides, and only a thread with a
      // remote parent pays for the owner lookup and the peer pointer.
      const long long own0 = (long long)pr.rank * pr.Nl, own1 = own0 + pr.Nl;
      bool all_local = true;
#pragma unroll
      for (int p = 0; p < PPT; ++p) all_local &= !valid[p] || (s[p] >= own0 && s[p] < own1);
      if (all_local) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) src[p] = xloc + (valid[p] ? s[p] - own0 : 0);
      } else {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          const unsigned g = (unsigned)s[p];
          const unsigned q = owner_of(pr, g);
          src[p] = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < PPT; ++p) src[p] = xloc + s[p];
    }
  }
  real g[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) g[p] = (real)0;

#pragma unroll
  for (int kk = 0; kk < d; kk += 4) {
    // issue the loads of this chunk of (up to) 4 coordinates first
    real xv[4][PPT];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kk + j;
#pragma unroll
      for (int p = 0; p < PPT; ++p) xv[j][p] = (k < d && valid[p]) ? (COH ? __ldcg(src[p] + (long long)k * Ns) : __ldg(src[p] + (long long)k * Ns)) : (real)0;
    }
    // noise: counter = (global slot, step, chunk)
    real z[4][PPT];
    if (D > 0 && zpre != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < PPT; ++p) z[j][p] = zpre[kk >> 2][j][p];
    } else if (zinj == nullptr) {
      chunk_noise<real, PPT>(d, kk, slot0, i0, step, key0, key1, z);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < PPT; ++p)
          z[j][p] = (kk + j < d && valid[p]) ? (real)zinj[(long long)(kk + j) * N + i0 + p] : (real)0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kk + j;
      if (k < d) {
        real A = a.A[k], Bk = a.D[k], S = a.S[k], Cc = a.C[k];
        real xn[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          xn[p] = r_fma<real>(S, z[j][p], r_fma<real>(A, xv[j][p], Bk));
          g[p] = r_fma<real>(Cc, xn[p], g[p]);
        }
        real* dst = xdst + (long long)k * Ns + i0;
        if (full) {
          vec_t v;
          real* vp = reinterpret_cast<real*>(&v);
#pragma unroll
          for (int p = 0; p < PPT; ++p) vp[p] = xn[p];
          *reinterpret_cast<vec_t*>(dst) = v;
        } else {
#pragma unroll
          for (int p = 0; p < PPT; ++p)
            if (valid[p]) dst[p] = xn[p];
        }
      }
    }
  }
  mx = -__longlong_as_double(0x7FF0000000000000ll);  // -inf
  bad = false;
  if (a.has_obs) {
    real lw[PPT];
    real mxr = (real)mx;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      lw[p] = obs_loglik<real>(a, g[p]);
      if (valid[p]) {
        if (lw[p] != lw[p]) bad = true;
        else mxr = lw[p] > mxr ? lw[p] : mxr;
      }
    }
    mx = (double)mxr;
    if (lw_out != nullptr) {
#pragma unroll
      for (int p = 0; p < PPT; ++p) lw_out[p] = lw[p];
    }
    if (full) {
      vec_t v;
      real* vp = reinterpret_cast<real*>(&v);
#pragma unroll
      for (int p = 0; p < PPT; ++p) vp[p] = lw[p];
      *reinterpret_cast<vec_t*>(logw + i0) = v;
    } else {
#pragma unroll
      for (int p = 0; p < PPT; ++p)
        if (valid[p]) logw[i0 + p] = lw[p];
    }
  }
}

#ifndef CSSM_K1_MINBLOCKS
#define CSSM_K1_MINBLOCKS 4
#endif
// chunks of 256 * PPT particles per block of a sharded filter's K1 (the host sizes the grid).  Measured on two ranks of 2^26
// particles: 1 chunk 0.841 ms, 4 chunks 0.796, 13 chunks (eight waves of blocks, run-time count) 0.817; unsharded 0.732
constexpr int K1_SH_CHUNKS = 4;
template <typename real, int D, int PPT = VecOf<real>::PPT, bool SH = false>
__global__ void __launch_bounds__(256, PPT == 4 ? CSSM_K1_MINBLOCKS : 6)
k_propagate_weight(const __grid_constant__ StepArgs<real> a, const __grid_constant__ Peers pr, real* __restrict__ xdst,
                   const int32_t* __restrict__ anc, real* __restrict__ logw, const double* __restrict__ zinj,
                   long long N, long long Ns, unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step,
                   K1Ctl ctl) {
  griddep_wait();
  griddep_launch();
  if (SH && pr.R > 1) {  // (sharded filters run k_propagate_weight_sh below; kept so that the single-rank code is what it was)
    const XchSlot* mine = pr.xch[pr.rank];
    gate_wait(&ctl.sc->gate1, ctl.gstep, pr, ctl.sc, [&](int q) { return &mine[q].progress; });
  }
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PPT;
  double mx;
  bool bad;
  if ((long long)(blockIdx.x + 1) * blockDim.x * PPT <= N)  // all but the last block: no per-particle bounds predicates
    propagate_particles<real, D, PPT, false, true, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mx, bad);
  else
    propagate_particles<real, D, PPT, false, false, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mx, bad);
  if (!a.has_obs && !(SH && pr.R > 1)) return;
  k1_tail<SH>(mx, bad, a.has_obs, pr, ctl);
}

// K1 of a sharded filter (its own kernel: the single-rank one above keeps the code the optimiser gives it alone)
template <typename real, int D, int PPT = VecOf<real>::PPT>
__global__ void __launch_bounds__(256, PPT == 4 ? CSSM_K1_MINBLOCKS : 6)
k_propagate_weight_sh(const __grid_constant__ StepArgs<real> a, const __grid_constant__ Peers pr, real* __restrict__ xdst,
                   const int32_t* __restrict__ anc, real* __restrict__ logw, const double* __restrict__ zinj,
                   long long N, long long Ns, unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step,
                   K1Ctl ctl) {
  constexpr bool SH = true;
  griddep_wait();
  griddep_launch();
  // Sharded: a block takes K1_SH_CHUNKS consecutive chunks of 256 * PPT particles.  What a block of a sharded filter does
  // once -- the test below with its two loads, and at the end a fence and a ticket atomic whose answer it has to wait for
  // -- costs about a microsecond, a sixth of the life of a one-chunk block.
  constexpr int CH = SH ? K1_SH_CHUNKS : 1;
  const long long blk0 = (long long)blockIdx.x * blockDim.x * PPT * CH;  // first particle of the block
  if (SH && pr.R > 1) {
    // The peers must have finished the previous step -- before a block reads ancestors a PEER's search wrote.  The
    // output slots [out_lo, out_hi) of the last resampling were written by this rank's own search (complete: it is the
    // preceding kernel of the stream), and with exchangeable particles that is all but the slots near the rank borders:
    // a block inside the range starts at once, and the rendezvous with the peers hides behind the kernel instead of
    // standing in front of it.  What else the rendezvous used to order is ordered without it: the cloud written here was
    // last read by the peers' K1 of the previous step (finished before this rank's search could pass its own gate),
    // log-weights are double-buffered by observed-step parity (a peer's walk over a run of repeated keys may still read
    // the previous buffer).  local_ok == 0, or the previous step had no observation: every block waits.
    if (ctl.pub_here && blockIdx.x == 0 && threadIdx.x < 32) push_progress_warp(pr, ctl.gstep);  // our kernels of the steps before are complete
    const long long g0 = (long long)pr.rank * pr.Nl + blk0, g1 = g0 + (long long)blockDim.x * PPT * CH;
    const bool interior = ctl.local_ok && anc != nullptr && g0 >= ctl.sc->out_lo && g1 <= ctl.sc->out_hi;  // block-uniform
    if (!interior) {
      const XchSlot* mine = pr.xch[pr.rank];
      gate_wait(&ctl.sc->gate1, ctl.gstep, pr, ctl.sc, [&](int q) { return &mine[q].progress; });
    }
  }
  double mx;
  bool bad;
  if (!SH) {
    const long long i0 = blk0 + (long long)threadIdx.x * PPT;
    if ((long long)(blockIdx.x + 1) * blockDim.x * PPT <= N)  // all but the last block: no per-particle bounds predicates
      propagate_particles<real, D, PPT, false, true, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mx, bad);
    else
      propagate_particles<real, D, PPT, false, false, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mx, bad);
  } else {
    mx = -__longlong_as_double(0x7FF0000000000000ll);
    bad = false;
#pragma unroll 1
    for (int c = 0; c < CH; ++c) {
      const long long c0 = blk0 + (long long)c * blockDim.x * PPT;
      if (c0 >= N) break;
      const long long i0 = c0 + (long long)threadIdx.x * PPT;
      double mxc;
      bool badc;
      if (c0 + (long long)blockDim.x * PPT <= N)
        propagate_particles<real, D, PPT, false, true, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mxc, badc);
      else
        propagate_particles<real, D, PPT, false, false, SH>(a, pr, xdst, anc, logw, zinj, N, Ns, slot0, key0, key1, step, i0, mxc, badc);
      mx = fmax(mx, mxc);
      bad |= badc;
    }
  }
  if (!a.has_obs && !(SH && pr.R > 1)) return;
  k1_tail<SH>(mx, bad, a.has_obs, pr, ctl);
}

// ---------------------------------------------------------------------------------------------
// K1' LGCP: n_sub sub-steps of the SDE's transition in registers, cumulative hazard, log-weight
//     logw = f(x_n, t) - sum_i exp(f(x_i, t_i)) * delta.  a.A/M/D/S are the constants of ONE
//     sub-step of length delta; ctab[s*d + k] are the f-coefficients at sub-step time t_i
//     (NULL: use a.C for every sub-step, i.e. no seasonal leaf); a.C are those at time t.
//     n_sub == 0 is dt == 0: state unchanged, log-weight f - f (model/ParticleFilter.scala:212-213).
// ---------------------------------------------------------------------------------------------
template <typename real, int DP>
__global__ void __launch_bounds__(256)
k_lgcp_weight(const __grid_constant__ StepArgs<real> a, const __grid_constant__ Peers pr, real* __restrict__ xdst,
              const int32_t* __restrict__ anc, real* __restrict__ logw, const double* __restrict__ zinj,
              const real* __restrict__ ctab, long long n_sub, real delta, long long N, long long Ns,
              unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step, K1Ctl ctl) {
  constexpr int PC = Normals<real>::PER_CALL;
  griddep_wait();
  griddep_launch();
  if (pr.R > 1) {
    if (ctl.pub_here && blockIdx.x == 0 && threadIdx.x < 32) push_progress_warp(pr, ctl.gstep);  // our kernels of the steps before are complete
    const XchSlot* mine = pr.xch[pr.rank];
    gate_wait(&ctl.sc->gate1, ctl.gstep, pr, ctl.sc, [&](int q) { return &mine[q].progress; });
  }
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  double mx = -__longlong_as_double(0x7FF0000000000000ll);
  bool bad = false;
  if (valid) {
    const real* src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
    if (anc) {
      if (pr.R > 1) {
        const unsigned g = (unsigned)anc[i];
        const unsigned q = owner_of(pr, g);
        src = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
      } else {
        src = reinterpret_cast<const real*>(pr.x[0]) + anc[i];
      }
    }
    real x[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) x[k] = (k < a.d) ? src[(long long)k * Ns] : (real)0;
    unsigned long long slot = slot0 + (unsigned long long)i;
    real hz = (real)0;
    // the normals of this particle and step form one stream n = s*d + k; Philox call n / PC yields
    // elements n % PC, so every generated normal is used (a d = 1 model takes one call per 4 sub-steps)
    real zb[PC];
    int zpos = PC;
    uint32_t ncall = 0;
    long long s_first = 0;
    if (DP == 1 && zinj == nullptr) {
      // One coordinate (the LGCP of BASELINE configs[2]): Philox call c feeds the sub-steps c*PC .. c*PC+PC-1, so
      // the loop runs call by call with the PC sub-steps unrolled -- no position counter, no select chain, no
      // branch per sub-step.  The same operations in the same order as the general loop below: same bits.
      const real A = a.A[0], Dc = a.D[0], S = a.S[0];
      real xx = x[0];
      const long long n_full = n_sub - (n_sub % PC);
      for (long long s = 0; s < n_full; s += PC) {
        Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | ncall, key0, key1, zb);
        ++ncall;
#pragma unroll
        for (int q = 0; q < PC; ++q) {
          xx = r_fma<real>(S, zb[q], r_fma<real>(A, xx, Dc));
          const real cc = ctab ? ctab[s + q] : a.C[0];
          hz += r_exp<real>(r_fma<real>(cc, xx, (real)0)) * delta;
        }
      }
      x[0] = xx;
      s_first = n_full;  // the last n_sub % PC sub-steps take the general loop (zpos == PC: it draws call `ncall`)
    }
    for (long long s = s_first; s < n_sub; ++s) {
      real gs = (real)0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        if (k < a.d) {
          real z;
          if (zinj == nullptr) {
            if (zpos == PC) {
              Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | ncall, key0, key1, zb);
              ++ncall;
              zpos = 0;
            }
            z = zb[0];
#pragma unroll
            for (int q = 1; q < PC; ++q) z = (zpos == q) ? zb[q] : z;
            ++zpos;
          } else {
            z = (real)zinj[((long long)s * a.d + k) * N + i];
          }
          x[k] = r_fma<real>(a.S[k], z, r_fma<real>(a.A[k], x[k], a.D[k]));
          real cc = ctab ? ctab[s * a.d + k] : a.C[k];
          gs = r_fma<real>(cc, x[k], gs);
        }
      }
      hz += r_exp<real>(gs) * delta;
    }
    real g = (real)0;
#pragma unroll
    for (int k = 0; k < DP; ++k)
      if (k < a.d) {
        g = r_fma<real>(a.C[k], x[k], g);
        xdst[(long long)k * Ns + i] = x[k];
      }
    real lw = (n_sub == 0) ? (g - g) : (g - hz);
    logw[i] = lw;
    if (lw != lw) bad = true;
    else mx = (double)lw;
  }
  k1_tail<true>(mx, bad, 1, pr, ctl);
}

// ---------------------------------------------------------------------------------------------
// weights: w1 = exp(logw - gmax) in the filter's own precision (F32 filters: expf_det on the fp32
// pipe and integer-only fixed point; F64 filters: exp_det), or a caller-given fp64 array
// (cssm_resample, real = double).  `wt` is the type a weight is held in; (double)w is exact.
// ---------------------------------------------------------------------------------------------
template <typename real> struct WeightSrc;
template <> struct WeightSrc<float> {
  typedef float wt;
  static constexpr int Q2 = 48;  // quantum of the sum of squares: 2^-48 (64-bit partial sums)
  const float* logw;
  const double* direct;  // never set for fp32 filters
  double gmax;           // the max of fp32 log-weights: exactly an fp32 value
  __device__ __forceinline__ float weight(float lw) const { return expf_det(__fsub_rn(lw, (float)gmax)); }
  __device__ __forceinline__ float operator()(long long idx) const { return weight(logw[idx]); }
  __device__ __forceinline__ static u128 fix(float w, int) { return fix_f32(w); }
  // sum of squares: quantum 2^-48, a whole block's partial sum stays below 2^64 -> 64-bit adds and shuffles
  __device__ __forceinline__ static void acc_sq(u128& a, float w, double) { a.lo += fix_sq48_f32(w); }
  __device__ __forceinline__ static u128 warp_sum_sq(u128 a) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) a.lo += __shfl_xor_sync(0xffffffffu, a.lo, m);
    return make_u128(a.lo, 0);
  }
  // non-negative floats order as their bit patterns: one REDUX instead of five shuffle rounds
  __device__ __forceinline__ static double warp_max(float w) {
    return (double)__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(w)));
  }
  // weights of ITEMS elements base, base+stride, ...: loads first, then the arithmetic
  template <int ITEMS>
  __device__ __forceinline__ void load(long long base, int stride, long long N, float* wv) const {
    float lw[ITEMS];
    if (base + (long long)(ITEMS - 1) * stride < N) {
      if (stride == 1 && ITEMS == 8 && (base & 3) == 0) {
        const float4 a = *reinterpret_cast<const float4*>(logw + base), b = *reinterpret_cast<const float4*>(logw + base + 4);
        lw[0] = a.x; lw[1] = a.y; lw[2 % ITEMS] = a.z; lw[3 % ITEMS] = a.w;
        lw[4 % ITEMS] = b.x; lw[5 % ITEMS] = b.y; lw[6 % ITEMS] = b.z; lw[7 % ITEMS] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) lw[j] = logw[base + (long long)j * stride];
      }
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) wv[j] = weight(lw[j]);
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const long long idx = base + (long long)j * stride;
        wv[j] = (idx < N) ? weight(logw[idx]) : 0.0f;
      }
    }
  }
};
template <> struct WeightSrc<double> {
  typedef double wt;
  static constexpr int Q2 = 96;
  const double* logw;    // in-filter source (NULL when `direct` is used)
  const double* direct;
  double gmax;
  __device__ __forceinline__ double weight(double raw) const { return direct ? raw : exp_det(__dsub_rn(raw, gmax)); }
  __device__ __forceinline__ double operator()(long long idx) const { return direct ? direct[idx] : weight(logw[idx]); }
  __device__ __forceinline__ static u128 fix(double w, int qb) { return fix_fast(w, qb); }
  // direct weights are pre-scaled to <= 1 by q2scale = 2^-(96-qb) before squaring
  __device__ __forceinline__ static void acc_sq(u128& a, double w, double q2scale) {
    const double ws = __dmul_rn(w, q2scale);
    a = add128(a, fix_fast(__dmul_rn(ws, ws), 96));
  }
  __device__ __forceinline__ static u128 warp_sum_sq(u128 a) { return warp_sum128(a); }
  __device__ __forceinline__ static double warp_max(double w) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, m));
    return w;
  }
  template <int ITEMS>
  __device__ __forceinline__ void load(long long base, int stride, long long N, double* wv) const {
    const double* src = direct ? direct : logw;
    double raw[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const long long idx = base + (long long)j * stride;
      raw[j] = (idx < N) ? src[idx] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const long long idx = base + (long long)j * stride;
      wv[j] = (idx < N) ? weight(raw[j]) : 0.0;
    }
  }
};

// max of a caller-given weight array -> acc[0].gmax_key (cssm_resample only)
static __global__ void __launch_bounds__(256) k_max_direct(const double* __restrict__ w, long long N, FilterScalars* sc) {
  double mx = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) mx = fmax(mx, w[i]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  if ((threadIdx.x & 31) == 0) atomicMax(&sc->acc[0].gmax_key, ord_key(mx));
}
// In-filter: gmax = the max log-weight, weights are <= 1, quantum 2^-96.  Direct weights:
// power-of-two pre-scale so that every weight is <= 1, qb = 96 - e with max = f * 2^e, f in [0.5, 1).
struct PreScan {
  double gmax;
  int qb;
};
__device__ __forceinline__ PreScan pre_scan(unsigned long long gmax_key, bool direct) {
  double mx = ord_unkey(gmax_key);
  PreScan p;
  p.gmax = direct ? 0.0 : mx;
  p.qb = 96;
  if (direct && mx > 1.0) {
    unsigned long long b = (unsigned long long)__double_as_longlong(mx);
    p.qb = 96 - ((int)((b >> 52) & 0x7ff) - 1022);
  }
  return p;
}

// per-filter tables of the weight pass
struct SumTables {
  u128* tile_sum;             // [nt]   exact sum of fix(w) over the tile
  u128* tile_q;               // [nt]   exact sum of fix(w^2) over the tile (flat mode, ns == 0: no super tiles, no atomics --
                              //        every block of K3 adds the tile sums itself; clouds of at most a few thousand tiles)
  double* tile_maxw;          // [nt]   max weight of the tile
  u128* super_sum;            // [2][ns] by observed-step parity, zeroed for the next step by K3
  u128* super_q;              // [2][ns]
  unsigned long long* super_ticket;  // [ns] monotone
  int nt, ns;
};

__device__ __forceinline__ u128 block_sum128(u128 v, u128* s_warp) {  // result valid in thread 0
  v = warp_sum128(v);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  u128 t = make_u128(0, 0);
  if (threadIdx.x == 0)
    for (int w = 0; w < TILE_THREADS / 32; ++w) t = add128(t, s_warp[w]);
  return t;
}
__device__ __forceinline__ u128 ld_gpu128(const u128* p) {
  return make_u128(ld_gpu(&p->lo), ld_gpu(&p->hi));
}
// exact sum / largest weight of tile tl of rank q, as another block published it (L2-coherent reads)
__device__ __forceinline__ u128 walk_tile_sum(const SumTables& tb, const Peers& pr, int q, int tl) {
  return ld_gpu128((q == pr.rank) ? &tb.tile_sum[tl] : &pr.tile_sum[q][tl]);
}
__device__ __forceinline__ double walk_tile_maxw(const SumTables& tb, const Peers& pr, int q, int tl) {
  return __longlong_as_double((long long)ld_gpu((const unsigned long long*)((q == pr.rank) ? &tb.tile_maxw[tl] : &pr.tile_maxw[q][tl])));
}

// K2  w1 = exp(logw - max); exact sums of w1 and w1^2 per tile, per super tile and per rank.
//     Strided access inside the tile: the order is irrelevant for an exact sum.
template <typename real, int ITEMS, bool SH = false>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_sums(const real* __restrict__ logw, const double* __restrict__ direct, long long N, FilterScalars* __restrict__ sc,
              int parity, unsigned long long obs_seq, SumTables tb, const __grid_constant__ Peers pr, int pub_here = 0) {
  constexpr int TILE = TILE_THREADS * ITEMS;
  __shared__ u128 s_w[TILE_THREADS / 32];
  __shared__ double s_mxw[TILE_THREADS / 32];
  __shared__ unsigned long long s_key;
  griddep_wait();
  griddep_launch();
  const int RK = SH ? pr.R : 1;
  StepAcc* A = &sc->acc[parity];
  if (RK > 1) {  // all-gather of the per-rank maxima: every rank pushes its own into our slots (block 0 of this kernel)
    if (pub_here && blockIdx.x == 0 && threadIdx.x < 32) publish_max_warp(pr, sc, parity, obs_seq);
    const XchSlot* mine = pr.xch[pr.rank];
    gate_wait(&sc->gate2, obs_seq + 1, pr, sc, [&](int q) { return &mine[q].max_seq[parity]; });
    if (threadIdx.x < 32) {
      unsigned long long key = ((int)threadIdx.x < RK) ? ld_relaxed_sys(&mine[threadIdx.x].max_key[parity]) : 0ull;
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, m);
        key = o > key ? o : key;
      }
      if (threadIdx.x == 0) s_key = key;
    }
  } else if (threadIdx.x == 0) {
    s_key = A->gmax_key;
  }
  __syncthreads();
  const PreScan ps = pre_scan(s_key, direct != nullptr);
  WeightSrc<real> ws{logw, direct, ps.gmax};
  typedef typename WeightSrc<real>::wt wt;
  const int qb = ps.qb;
  const double q2scale = __longlong_as_double((long long)(1023 - (96 - qb)) << 52);  // direct weights: w * 2^-(96-qb) <= 1
  const long long base = (long long)blockIdx.x * TILE;
  u128 acc = make_u128(0, 0), acc2 = make_u128(0, 0);
  wt wv[ITEMS];
  ws.template load<ITEMS>(base + threadIdx.x, TILE_THREADS, N, wv);
  wt mxv = (wt)0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    acc = add128(acc, WeightSrc<real>::fix(wv[j], qb));
    WeightSrc<real>::acc_sq(acc2, wv[j], q2scale);
    mxv = wv[j] > mxv ? wv[j] : mxv;
  }
  const double mxw = WeightSrc<real>::warp_max(mxv);
  acc2 = WeightSrc<real>::warp_sum_sq(acc2);
  __shared__ u128 s_w2[TILE_THREADS / 32];
  if ((threadIdx.x & 31) == 0) {
    s_w2[threadIdx.x >> 5] = acc2;
    s_mxw[threadIdx.x >> 5] = mxw;
  }
  const u128 t = block_sum128(acc, s_w);  // contains the __syncthreads that publishes s_w2 / s_mxw
  if (threadIdx.x == 0) {
    u128 t2 = s_w2[0];
    double m2 = s_mxw[0];
    for (int w = 1; w < TILE_THREADS / 32; ++w) {
      t2 = add128(t2, s_w2[w]);
      m2 = fmax(m2, s_mxw[w]);
    }
    tb.tile_sum[blockIdx.x] = t;
    tb.tile_maxw[blockIdx.x] = m2;
    if (tb.ns == 0) {  // flat mode: the tile sums are all K3 needs, no atomics and no dependent round trips at the tail
      tb.tile_q[blockIdx.x] = t2;
      return;
    }
    const int sidx = blockIdx.x / SUPER;
    u128* ssum = tb.super_sum + (size_t)parity * tb.ns + sidx;
    u128* ssq = tb.super_q + (size_t)parity * tb.ns + sidx;
    atomic_add128(ssum, t);
    atomic_add128(ssq, t2);
    __threadfence();
    const unsigned long long in_super = (unsigned long long)min(SUPER, tb.nt - sidx * SUPER);
    const unsigned long long tk = atomicAdd(&tb.super_ticket[sidx], 1ull);
    if (tk % in_super == in_super - 1) {  // last tile of this super tile: fold it into the rank totals
      __threadfence();
      atomic_add128(&A->tot, ld_gpu128(ssum));
      atomic_add128(&A->q, ld_gpu128(ssq));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3  systematic / stratified: inclusive CDF of a tile in shared memory + ancestor search
// ---------------------------------------------------------------------------------------------
// k_i of model/Resampling.scala:69 (systematic) and :82-83 (stratified), times the total.  KIND is a
// template parameter so that the systematic kernel carries no Philox code at all.  When n is a
// power of two the division by n is the (exact) multiplication by inv_n -- same bits, no DDIV call.
template <int KIND>
struct KFun {
  double u, n, inv_n, total;  // inv_n == 0: n is not a power of two
  const double* uarr;         // stratified, injected
  uint32_t key0, key1, step;
  __device__ __forceinline__ double ui(long long i) const {
    if (uarr) return uarr[i];
    uint4 v = philox4x32(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), step, RNG_RESAMPLE | 1u), key0, key1);
    return u64_to_unit_double(v.x, v.y);
  }
  __device__ __forceinline__ double k(long long i) const {
    const double num = (KIND == CSSM_RESAMPLE_SYSTEMATIC) ? __dadd_rn(u, (double)i) : __dadd_rn((double)i, ui(i));
    return (inv_n != 0.0) ? __dmul_rn(num, inv_n) : __ddiv_rn(num, n);
  }
  // the key of output i in the un-normalised domain: C_j >= k_i  <=>  P_j >= k_i * total
  __device__ __forceinline__ double operator()(long long i) const { return __dmul_rn(k(i), total); }
  // count_le(P) by arithmetic.  The keys are (nearly) an arithmetic progression, so the count is
  // floor(P*n/total - u) + 1 (systematic) or floor(P*n/total) plus one exactly evaluated key
  // (stratified) -- unless the real number being floored is within 1e-5 of an integer: the rounding
  // of the keys and of this estimate moves the boundary by less than 1.5e-6 for n <= 2^31, so
  // outside that band the result is provably the exact count, inside it the exact loop decides.
  // scale = fl(n / total).
  __device__ __forceinline__ long long count_fast(double P, double scale, long long N) const {
    const double x = (KIND == CSSM_RESAMPLE_SYSTEMATIC) ? __dsub_rn(__dmul_rn(P, scale), u) : __dmul_rn(P, scale);
    const double fl = floor(x);
    const double fr = __dsub_rn(x, fl);
    if (!(fr >= 1e-5 && fr <= 1.0 - 1e-5) || !(x < 4.0e9) || !(x > -4.0e9)) return count_le(P, N);
    long long i0 = (long long)fl;
    if (KIND == CSSM_RESAMPLE_SYSTEMATIC) {
      i0 += 1;
    } else {
      if (i0 >= 0 && i0 < N) i0 += ((*this)(i0) <= P) ? 1 : 0;
    }
    return i0 < 0 ? 0 : (i0 > N ? N : i0);
  }
  // the same count minus `lo` (as a double, lo <= count), clamped to [0, n_rel]: 32-bit arithmetic
  __device__ __forceinline__ int count_rel(double P, double scale, long long N, long long lo, double lo_d, int n_rel) const {
    if (KIND == CSSM_RESAMPLE_SYSTEMATIC) {
      // y = P*scale - (u + lo) in one fma; the count minus lo is floor(y) + 1 unless y is within 1e-5 of an
      // integer, which is the case exactly when the floors of y - 1e-5 and y + 1e-5 differ.  The conversions
      // saturate, the clamp below absorbs that.  (u + lo is the same for every particle of the thread.)
      const double y = __fma_rn(P, scale, -__dadd_rn(u, lo_d));
      const int a = __double2int_rd(__dadd_rn(y, -1e-5)), b = __double2int_rd(__dadd_rn(y, 1e-5));
      int c;
      if (a != b) {
        const long long d = count_le(P, N) - lo;
        c = (int)(d > 0x7FFFFFFFll ? 0x7FFFFFFFll : (d < -0x7FFFFFFFll ? -0x7FFFFFFFll : d));
      } else {
        c = (a == 0x7FFFFFFF) ? a : a + 1;
      }
      return min(max(c, 0), n_rel);
    }
    const double x = (KIND == CSSM_RESAMPLE_SYSTEMATIC) ? __dsub_rn(__dmul_rn(P, scale), u) : __dmul_rn(P, scale);
    const double fl = floor(x);
    const double fr = __dsub_rn(x, fl);
    int c;
    if (!(fr >= 1e-5 && fr <= 1.0 - 1e-5) || !(x < 4.0e9) || !(x > -4.0e9)) {
      const long long d = count_le(P, N) - lo;
      c = (int)(d > 0x7FFFFFFFll ? 0x7FFFFFFFll : d);
    } else {
      c = __double2int_rn(__dsub_rn(fl, lo_d));  // exact integers; the conversion saturates
      if (KIND == CSSM_RESAMPLE_SYSTEMATIC) {
        c = (c == 0x7FFFFFFF) ? c : c + 1;
      } else {
        const long long i0 = (long long)fl;
        if (i0 >= 0 && i0 < N && (*this)(i0) <= P) c = (c == 0x7FFFFFFF) ? c : c + 1;
      }
    }
    return min(max(c, 0), n_rel);
  }
  // number of outputs i in [0, N) with key_i <= c   (key_i is non-decreasing in i); exact, by evaluation
  __device__ __noinline__ long long count_le(double c, long long N) const {
    if (!(c >= 0.0)) return 0;
    double est = c / total * n - (KIND == CSSM_RESAMPLE_SYSTEMATIC ? u : 0.0);
    long long i = (est >= (double)N) ? N - 1 : (long long)floor(est);
    if (i < 0) i = 0;
    if (i > N - 1) i = N - 1;
    while (i + 1 < N && (*this)(i + 1) <= c) ++i;  // a few steps at most
    while (i >= 0 && (*this)(i) > c) --i;
    return i + 1;
  }
};

// shared-memory index of element i of a tile.  A thread owns ITEMS consecutive elements; with 8
// doubles per thread the warp's stores would all fall into two bank groups, one pad double per 8
// elements spreads them over all banks.
template <int ITEMS> __device__ __forceinline__ int phys(int i) { return ITEMS == 8 ? i + (i >> 3) : i; }
template <int ITEMS> struct TileSmem { static constexpr int SIZE = TILE_THREADS * ITEMS + (ITEMS == 8 ? TILE_THREADS : 0); };

__device__ __forceinline__ float to_float_rd(float w) { return w; }
__device__ __forceinline__ float to_float_rd(double w) { return __double2float_rd(w); }

// inclusive CDF values P_j = dbl128(exact prefix) of one tile into Ps[phys(0..TILE)) and the
// weights (as double) into Ws; `excl` = exact sum of everything before the tile.  All threads call.
template <typename real, int ITEMS>
__device__ __forceinline__ void tile_cdf(const WeightSrc<real>& ws, int qb, u128 excl, long long tile0, long long N,
                                         double* Ps, double* Ws, u128* s_warp, double* Pv = nullptr,
                                         const typename WeightSrc<real>::wt* wv_in = nullptr, float* minw_out = nullptr) {
  typedef typename WeightSrc<real>::wt wt;
  const long long base = tile0 + (long long)threadIdx.x * ITEMS;
  wt wv[ITEMS];
  if (wv_in != nullptr) {  // the caller already holds this thread's ITEMS weights (same values, same order)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) wv[j] = wv_in[j];
  } else {
    ws.template load<ITEMS>(base, 1, N, wv);
  }
  if (minw_out != nullptr) {  // a lower bound (fp32, rounded down) of this thread's smallest weight
    float m = 3.4028234663852886e38f;
    if (base + ITEMS <= N) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) m = fminf(m, to_float_rd(wv[j]));
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j)
        if (base + j < N) m = fminf(m, to_float_rd(wv[j]));
    }
    *minw_out = m;
  }
  u128 e[ITEMS];
  u128 run = make_u128(0, 0);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (Ws) Ws[phys<ITEMS>(threadIdx.x * ITEMS + j)] = (double)wv[j];
    run = add128(run, WeightSrc<real>::fix(wv[j], qb));
    e[j] = run;
  }
  // exclusive scan of the thread totals across the block
  u128 incl = run;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u128 o = shfl_up128(incl, d);
    if (lane >= d) incl = add128(incl, o);
  }
  __syncthreads();  // s_warp may still be read by the caller's previous use
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  u128 off = excl;
  for (int w = 0; w < wid; ++w) off = add128(off, s_warp[w]);
  u128 ex = shfl_up128(incl, 1);
  if (lane == 0) ex = make_u128(0, 0);
  off = add128(off, ex);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const double P = dbl128(add128(off, e[j]), qb);
    Ps[phys<ITEMS>(threadIdx.x * ITEMS + j)] = P;
    if (Pv) Pv[j] = P;
  }
  __syncthreads();
}

// does adding the next weight leave the cumulative value unchanged?  This is what makes a TreeMap
// key repeat in the reference (model/Resampling.scala:55-57), tested in the reference's normalised
// domain: C = fl(P/total), wn = fl(w/total).  A weight above 2^-52 * P cannot vanish (cheap filter).
static __device__ __noinline__ bool vanishes_exact(double P, double w, double total) {
  const double c = __ddiv_rn(P, total);
  return __dadd_rn(c, __ddiv_rn(w, total)) == c;
}
__device__ __forceinline__ bool vanishes(double P, double w, double total) {
  if (w > P * 2.220446049250313e-16) return false;
  // w <= 2^-55 P: fl(w/total) < 2^-54.9 fl(P/total) is below half an ulp of fl(P/total), the sum rounds
  // back to it -- decided without the two divisions (only the three binades in between need them: out of line)
  if (w <= P * 2.7755575615628914e-17) return true;
  return vanishes_exact(P, w, total);
}

struct K3Ctl {
  int parity;
  unsigned long long obs_seq, gstep;
  double inv_n;         // 1 / (number of outputs) when that is a power of two, else 0
  int direct;           // cssm_resample: caller weights, no ll update
  int add_ll, use_u_inj;
  int tie_first;        // CSSM_TIE_FIRST: plain inverse CDF (first index with C_j >= k), no TreeMap duplicate-key rule
  int defer_ll;         // the caller updates ll / ESS itself (ll_ess_update), off the critical path of the search
  int fast_ok;          // k_scan_search may try the certified fp64 path (cssm_filter_scan_mode; 0: exact path only)
  uint32_t key0, key1, step;
  double* ll_steps;
  int* ess_steps;
  long long step_slot;
  // single-launch series kernel (k_series_small): ancestors leave as 64-bit words (tag << 32 | index) that the consumer
  // polls -- no grid barrier between the search and the next gather.  NULL: plain int32 ancestors.
  unsigned long long* anc64;
  unsigned anc_tag;
  unsigned long long* dbg;  // CSSM_SERIES_DEBUG: shared-memory cycle counters of thread 0 ([7] = last stamp, [8..15] = stages of the search)
};
#define K3_STAMP(slot)                                                        \
  if (!PROTO3 && ctl.dbg != nullptr && threadIdx.x == 0) {                    \
    const unsigned long long now_ = (unsigned long long)clock64();            \
    ctl.dbg[slot] += now_ - ctl.dbg[7];                                       \
    ctl.dbg[7] = now_;                                                        \
  }

// ONE thread of the whole filter, once per observed step: ll += max + log(mean w1) (model/ParticleFilter.scala:127), ESS =
// floor(1 / sum wn^2) (:431-434) from the exact sums; PROTO3: also zero the accumulators of the next observed step.
template <typename real, bool PROTO3>
__device__ __noinline__ void ll_ess_update(FilterScalars* __restrict__ sc, const K3Ctl& ctl, u128 tot, u128 qsum, unsigned long long key,
                                           long long Ng, bool direct) {
  const PreScan ps = pre_scan(key, direct);
  const int qb = ps.qb;
  const double total = dbl128(tot, qb);
  const double gmax = ps.gmax;
  double incr = gmax + log(total / (double)Ng);
  int flags = 0;
  if (!(total > 0.0) || gmax != gmax || gmax - gmax != 0.0) {  // all weights zero / NaN / infinite max
    incr = __longlong_as_double(0x7FF8000000000000ll);
    flags |= FLAG_ZERO_TOTAL;
  }
  // sum (w/total)^2 = (exact sum w^2) / total^2; direct weights were pre-scaled by 2^-(96-qb)
  const double tsc = __dmul_rn(total, __longlong_as_double((long long)(1023 - (96 - qb)) << 52));
  const double s2 = __ddiv_rn(dbl128(qsum, WeightSrc<real>::Q2), __dmul_rn(tsc, tsc));
  const double inv = floor(1.0 / s2);
  const int ess = (inv == inv && inv < 2147483647.0) ? (int)inv : (inv == inv ? 2147483647 : 0);  // Scala .toInt saturates, NaN -> 0
  sc->gmax = gmax;
  sc->total = total;
  sc->qb = qb;
  sc->ll_incr = incr;
  if (ctl.add_ll) {
    const double ll = sc->ll + incr;
    sc->ll = ll;
    sc->ess = ess;
    if (ctl.ll_steps) ctl.ll_steps[ctl.step_slot] = ll;
    if (ctl.ess_steps) ctl.ess_steps[ctl.step_slot] = ess;
  }
  if (flags) atomicOr(&sc->flags, flags);
  if (PROTO3) {  // three-launch protocol: zero the accumulators of the next observed step
    StepAcc* nx = &sc->acc[ctl.parity ^ 1];
    nx->gmax_key = 0ull;
    nx->tot = make_u128(0, 0);
    nx->q = make_u128(0, 0);
  }
}

// the terms of ll_ess_update for one observed step, without touching the filter's scalars (k_series_one evaluates all
// steps at the end of its launch, one thread per step, and adds the increments in order): same expressions, same bits
template <typename real>
__device__ __forceinline__ void ll_ess_terms(u128 tot, u128 qsum, unsigned long long key, long long Ng, double& incr, int& ess, int& flags) {
  const PreScan ps = pre_scan(key, false);
  const int qb = ps.qb;
  const double total = dbl128(tot, qb);
  const double gmax = ps.gmax;
  incr = gmax + log(total / (double)Ng);
  flags = 0;
  if (!(total > 0.0) || gmax != gmax || gmax - gmax != 0.0) {
    incr = __longlong_as_double(0x7FF8000000000000ll);
    flags |= FLAG_ZERO_TOTAL;
  }
  const double tsc = __dmul_rn(total, __longlong_as_double((long long)(1023 - (96 - qb)) << 52));
  const double s2 = __ddiv_rn(dbl128(qsum, WeightSrc<real>::Q2), __dmul_rn(tsc, tsc));
  const double inv = floor(1.0 / s2);
  ess = (inv == inv && inv < 2147483647.0) ? (int)inv : (inv == inv ? 2147483647 : 0);
}

// ---- block-wide scan + search: cumulative values and weights of the tile in shared memory, expansion by head scatter +
//      block max-scan.  The kernel of the three-launch step (k_scan_search); measured faster there than the
//      warp-synchronous variant below (0.144 vs 0.166 ms at 2^24 particles), which the single-launch series kernels use. ----
// cdf_out == NULL: search (systematic / stratified), writes ancestors;
// cdf_out != NULL: write the un-normalised CDF (multinomial), no search
// shared memory of the scan + search of one tile
template <int ITEMS>
struct K3SmemBlk {
  static constexpr int TILE = TILE_THREADS * ITEMS;
  static constexpr int WIN = TILE + TILE_THREADS;  // outputs staged per pass (a tile has ~TILE offspring)
  double Ps[TileSmem<ITEMS>::SIZE];
  double Ws[TileSmem<ITEMS>::SIZE];
  int32_t s_res[WIN];
  u128 s_warp[TILE_THREADS / 32];
  u128 s_excl, s_tot, s_q, s_run;
  unsigned long long s_key;
  int s_cnt[TILE_THREADS], s_cnt2[TILE_THREADS / 32];
  long long s_pend, s_jfinal;
  double s_wnext, s_u, s_scale;
  int s_tp, s_brk;
  unsigned s_minw[TILE_THREADS / 32];  // per-warp min weight of the tile (fp32 bits; non-negative floats order as integers)
  int s_novanish;
};

// Tile t of the scan + search once the exact sums are known: `tot` / `qsum` = sum of fix(w1) / of
// fix(w1^2) over the whole filter, `key` = ordered key of the max log-weight, `excl` = exact sum of
// everything before the tile.  Block 0 also updates ll and ESS.  PROTO3: called from the
// three-launch step (zeroes the accumulators of the next observed step); the single-launch series
// kernel keeps its own.  All threads of the block call; returns whether a peer's memory was written.
template <typename real, int ITEMS, int KIND, bool PROTO3, bool SH = false>
__device__ __forceinline__ bool k3_tile_blk(K3SmemBlk<ITEMS>& sm, const real* __restrict__ logw, const double* __restrict__ direct,
                                        long long N, FilterScalars* __restrict__ sc, const SumTables& tb, const Peers& pr,
                                        const K3Ctl& ctl, const double* __restrict__ uarr, double* __restrict__ cdf_out, int t,
                                        u128 tot, u128 qsum, unsigned long long key, u128 excl,
                                        const typename WeightSrc<real>::wt* wv_in = nullptr) {
  constexpr int TILE = TILE_THREADS * ITEMS;
  const int RK = SH ? pr.R : 1, RNK = SH ? pr.rank : 0;  // single-rank instantiations carry no sharding code
  constexpr int WIN = K3SmemBlk<ITEMS>::WIN;
  double* Ps = sm.Ps;
  double* Ws = sm.Ws;
  int32_t* s_res = sm.s_res;
  u128* s_warp = sm.s_warp;
  int* s_cnt = sm.s_cnt;
  int* s_cnt2 = sm.s_cnt2;
  long long& s_pend = sm.s_pend;
  long long& s_jfinal = sm.s_jfinal;
  double& s_wnext = sm.s_wnext;
  double& s_u = sm.s_u;
  double& s_scale = sm.s_scale;
  u128& s_run = sm.s_run;
  int& s_tp = sm.s_tp;
  int& s_brk = sm.s_brk;
  const int nt = tb.nt;
  const long long Ng = (long long)RK * N;  // outputs of the whole (possibly sharded) filter
  const PreScan ps = pre_scan(key, direct != nullptr);
  const int qb = ps.qb;
  const double total = dbl128(tot, qb);

  // ---- block 0: ll increment max + log(mean w1), ESS = floor(1/sum wn^2), the resampling uniform
  //      is derived by every block; zero the accumulators of the next observed step ----------------
  if (threadIdx.x == 0) {
    double u;
    if (ctl.use_u_inj) {
      u = sc->u_inj;
    } else {
      uint4 v = philox4x32(make_uint4(0u, 0u, ctl.step, RNG_RESAMPLE), ctl.key0, ctl.key1);
      u = u64_to_unit_double(v.x, v.y);
    }
    s_u = u;
    s_scale = __ddiv_rn((double)Ng, total);
    s_pend = 0x7FFFFFFFFFFFFFFFll;
  }
  if (threadIdx.x == 64 && t == 0 && !ctl.defer_ll)  // another warp than the one deriving the uniform: the two run side by side
    ll_ess_update<real, PROTO3>(sc, ctl, tot, qsum, key, Ng, direct != nullptr);
  if (PROTO3 && t < tb.ns && threadIdx.x == 1) {
    tb.super_sum[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
    tb.super_q[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
  }

  WeightSrc<real> ws{logw, direct, ps.gmax};
  const long long tile0 = (long long)t * TILE;
  const int tile_n = (int)min((long long)TILE, N - tile0);
  if (threadIdx.x == 32 && cdf_out == nullptr) {
    // first weight after this tile (next tile, possibly the next rank's first particle); loaded here so
    // that its latency hides behind the tile scan
    double wn = 0.0;
    if (t < nt - 1) wn = (double)ws(tile0 + TILE);
    else if (RNK < RK - 1) wn = (double)WeightSrc<real>{reinterpret_cast<const real*>(pr.logw[RNK + 1]), nullptr, ps.gmax}(0);
    s_wnext = wn;
  }
  double Pv[ITEMS];
  float minw_thread;
  tile_cdf<real, ITEMS>(ws, qb, excl, tile0, N, Ps, cdf_out ? nullptr : Ws, s_warp, Pv, wv_in, &minw_thread);
  {
    const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(minw_thread));
    if ((threadIdx.x & 31) == 0) sm.s_minw[threadIdx.x >> 5] = mb;  // read after the next barrier
  }

  if (cdf_out != nullptr) {
    for (int j = threadIdx.x; j < tile_n; j += TILE_THREADS) cdf_out[tile0 + j] = Ps[phys<ITEMS>(j)];
    return false;
  }

  bool wrote_remote = false;
  // all weights zero / NaN (the reference divides by a zero total here and fails later): keep every
  // particle as its own ancestor; FLAG_ZERO_TOTAL is already raised
  const bool usable = (total > 0.0) && (total - total == 0.0);
  if (!usable) {
    for (int j = threadIdx.x; j < tile_n; j += TILE_THREADS) pr.anc[RNK][tile0 + j] = (int32_t)((long long)RNK * N + tile0 + j);
  } else {
  KFun<KIND> kf{s_u, (double)Ng, ctl.inv_n, total, uarr, ctl.key0, ctl.key1, ctl.step};
  const double c_end = Ps[phys<ITEMS>(tile_n - 1)];
  const bool last_tile = (t == nt - 1) && (RNK == RK - 1);
  const long long gbase = (long long)RNK * N + tile0;  // global index of the tile's first particle
  // ---- offspring counts: c_j = #{outputs with key <= P_j}; particle j owns the outputs [c_{j-1}, c_j) ----
  const double scale = s_scale;
  const long long lo = (t == 0 && RNK == 0) ? 0 : kf.count_fast(dbl128(excl, qb), scale, Ng);
  const double lo_d = (double)lo;
  const int n_rel = (int)(Ng - lo);
  int cr[ITEMS];  // counts relative to lo
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int idx = threadIdx.x * ITEMS + j;
    cr[j] = (last_tile && idx >= tile_n - 1) ? n_rel : kf.count_rel(Pv[j], scale, Ng, lo, lo_d, n_rel);
  }
  s_cnt[threadIdx.x] = cr[ITEMS - 1];
  __syncthreads();
  if (threadIdx.x == 0) {
    // A key can only repeat where a weight is at most 2^-52 of the cumulative value before it
    // (vanishes()).  If even the smallest weight of the tile is above 2^-52 of the tile's LAST
    // cumulative value, no key of this tile repeats and the per-output check is skipped.
    unsigned mm = sm.s_minw[0];
#pragma unroll
    for (int w = 1; w < TILE_THREADS / 32; ++w) mm = min(mm, sm.s_minw[w]);
    sm.s_novanish = (ctl.tie_first || (double)__uint_as_float(mm) > c_end * 2.220446049250313e-16) ? 1 : 0;
  }
  const int n_out = s_cnt[TILE_THREADS - 1];
  const long long hi = lo + n_out;
  // does the run of repeated keys at the end of this tile continue into the next tile?
  const bool cont = !ctl.tie_first && !last_tile && vanishes(c_end, s_wnext, total);
  if (last_tile && threadIdx.x == 0 && kf(Ng - 1) > c_end) atomicOr(&sc->flags, FLAG_CLAMPED);  // reference would throw (m.head)

  // ---- expansion, WIN outputs per pass: every particle with offspring drops its local index at the
  //      head of its range, an inclusive max-scan (indices grow with the position) fills the ranges,
  //      and the pass leaves as coalesced stores.  A heavy particle simply spans many passes. ----------
  {
    constexpr int PER = WIN / TILE_THREADS;  // outputs per thread and pass
    const int prev0 = threadIdx.x ? s_cnt[threadIdx.x - 1] : 0;
    int carry = 0;  // local index of the particle that owns the last output of the previous pass
    // the end of the run of repeated keys behind particle memo_j, as this thread last walked it: a heavy particle owns
    // every output of many passes, and each of them would otherwise walk the same run again (2^24 outputs x a run of a
    // thousand particles x two divisions per test: seconds)
    int memo_j = -1, memo_e = -1;
    constexpr int HEAVY_SPAN = 8 * WIN;  // a stretch of outputs this long without a head is filled directly
    for (int w0 = 0; w0 < n_out; w0 += WIN) {
      if (w0 > 0 && n_out - w0 >= HEAVY_SPAN) {  // block-uniform, and false for every tile of an ordinary cloud
        // One particle with very many offspring: where does the next particle's range begin?  Up to there every output
        // belongs to `carry`, the owner of the last output of the pass before -- no staging, no scan, plain stores.
        // (Window by window the 2^24 offspring of a single particle took 33 ms.)
        int nh = 0x7FFFFFFF;
        {
          int prev = prev0;
#pragma unroll
          for (int j = 0; j < ITEMS; ++j) {
            if (cr[j] > prev && prev >= w0) nh = min(nh, prev);
            prev = cr[j];
          }
        }
        nh = __reduce_min_sync(0xffffffffu, nh);
        if ((threadIdx.x & 31) == 0) s_cnt2[threadIdx.x >> 5] = nh;
        __syncthreads();
#pragma unroll
        for (int ww = 0; ww < TILE_THREADS / 32; ++ww) nh = min(nh, s_cnt2[ww]);
        __syncthreads();  // s_cnt2 is used again by the pass below
        const int span_end = min(nh, n_out);
        if (span_end - w0 >= HEAVY_SPAN) {
          int jt = carry;
          if (sm.s_novanish == 0 && jt + 1 < tile_n && !(Ws[phys<ITEMS>(jt + 1)] > Ps[phys<ITEMS>(jt)] * 2.220446049250313e-16)) {
            if (jt == memo_j) {
              jt = memo_e;
            } else {
              memo_j = jt;
              while (jt + 1 < tile_n && vanishes(Ps[phys<ITEMS>(jt)], Ws[phys<ITEMS>(jt + 1)], total)) ++jt;
              memo_e = jt;
            }
          }
          if (cont && jt == tile_n - 1 && threadIdx.x == 0) atomicMin(&s_pend, lo + w0);
          const int32_t val = (int32_t)(gbase + jt);
          const bool span_local = !SH || RK == 1 || (lo + w0 >= (long long)RNK * N && lo + span_end <= (long long)(RNK + 1) * N);
          if (span_local) {
            int32_t* const out = pr.anc[RNK] + (lo - (long long)RNK * N);
            for (int o = w0 + (int)threadIdx.x; o < span_end; o += TILE_THREADS) out[o] = val;
          } else {
            for (int o = w0 + (int)threadIdx.x; o < span_end; o += TILE_THREADS) {
              const long long i = lo + o;
              const unsigned q = owner_of(pr, (unsigned)i);
              pr.anc[q][i - (long long)q * N] = val;
              wrote_remote |= (q != RNK);
            }
          }
          w0 = span_end - WIN;  // the loop adds WIN: the next pass starts at the next particle's first output
          continue;
        }
      }
#pragma unroll
      for (int k = 0; k < PER; ++k) s_res[threadIdx.x * PER + k] = -1;
      __syncthreads();
      {
        int prev = prev0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          if (cr[j] > prev && prev >= w0 && prev < w0 + WIN) s_res[prev - w0] = threadIdx.x * ITEMS + j;
          prev = cr[j];
        }
      }
      __syncthreads();
      // inclusive max-scan of s_res: thread-local, then across the block
      int v[PER];
      int run = -1;
#pragma unroll
      for (int k = 0; k < PER; ++k) { run = max(run, s_res[threadIdx.x * PER + k]); v[k] = run; }
      int incl = run;
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = max(incl, o); }
      if (lane == 31) s_cnt2[wid] = incl;
      __syncthreads();
      int before = carry;
      for (int ww = 0; ww < wid; ++ww) before = max(before, s_cnt2[ww]);
      { const int o = __shfl_up_sync(0xffffffffu, incl, 1); if (lane > 0) before = max(before, o); }
      carry = max(carry, s_cnt2[TILE_THREADS / 32 - 1]);
      for (int ww = 0; ww < TILE_THREADS / 32 - 1; ++ww) carry = max(carry, s_cnt2[ww]);
#pragma unroll
      for (int k = 0; k < PER; ++k) s_res[threadIdx.x * PER + k] = max(v[k], before);
      __syncthreads();
      // copy-out: the duplicate-key rule per output, then a coalesced store
      const bool novanish = sm.s_novanish != 0;  // written before the barriers of this pass
      const int n_w = min(WIN, n_out - w0);
      int jts[PER];
      // all the shared-memory reads of the thread's PER outputs first, then the (rare) slow path, then the stores
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int o = threadIdx.x + k * TILE_THREADS;
        int jt = s_res[min(o, WIN - 1)];
        jt = (o < n_w) ? jt : 0;
        bool mayv = false;
        if (!novanish) {
          // TreeMap: a duplicated key keeps the last particle inserted.  First the cheap necessary condition
          // (next weight at most 2^-52 of the cumulative value), branch-free
          const int jn = min(jt + 1, tile_n - 1);
          mayv = (o < n_w) && (jt + 1 < tile_n) && !(Ws[phys<ITEMS>(jn)] > Ps[phys<ITEMS>(jt)] * 2.220446049250313e-16);
        }
        if (mayv) {
          if (jt == memo_j) {
            jt = memo_e;
          } else {
            memo_j = jt;
            while (jt + 1 < tile_n && vanishes(Ps[phys<ITEMS>(jt)], Ws[phys<ITEMS>(jt + 1)], total)) ++jt;
            memo_e = jt;
          }
        }
        jts[k] = jt;
      }
      int32_t* const out_local = pr.anc[RNK] + (lo + w0 - (long long)RNK * N);  // R == 1: plain coalesced stores
      // sharded: the outputs of a pass are consecutive slots; unless the pass straddles a rank border they all belong to
      // this rank (the common case: offspring stay near their parents) and take the same plain stores
      const bool pass_local = !SH || RK == 1 || (lo + w0 >= (long long)RNK * N && lo + w0 + n_w <= (long long)(RNK + 1) * N);
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int o = threadIdx.x + k * TILE_THREADS;
        if (o < n_w) {
          const int jt = jts[k];
          const long long i = lo + w0 + o;
          if (cont && jt == tile_n - 1) atomicMin(&s_pend, i);
          const int32_t val = (int32_t)(gbase + jt);
          if (!pass_local) {  // offspring slot i belongs to rank i / N: scatter over NVLink
            const unsigned q = owner_of(pr, (unsigned)i);
            pr.anc[q][i - (long long)q * N] = val;
            wrote_remote |= (q != RNK);
          } else {
            out_local[o] = val;
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const long long pend = s_pend;
  if (pend < hi) {
    // The selected run of repeated keys continues past this tile; its last element is the ancestor.
    // Walk forward over the GLOBAL tile sequence (rank-major): whole tiles are skipped from the
    // tables when every weight in them is strictly below half an ulp of the running (normalised)
    // value and the value stays in its binade, otherwise the tile is recomputed.
    const long long gnt = (long long)RK * nt;
    if (threadIdx.x == 0) { s_tp = RNK * nt + t + 1; s_run = add128(excl, tb.tile_sum[t]); s_jfinal = -1; }
    __syncthreads();
    for (;;) {
      if (threadIdx.x == 0) {
        long long tp = s_tp;
        u128 run = s_run;
        while (tp < gnt) {
          const int q = (int)(tp / nt), tl = (int)(tp % nt);
          const u128 tsum = ld_gpu128((q == RNK) ? &tb.tile_sum[tl] : &pr.tile_sum[q][tl]);  // L2: another block wrote it
          const double mxw = __longlong_as_double((long long)ld_gpu((const unsigned long long*)((q == RNK) ? &tb.tile_maxw[tl] : &pr.tile_maxw[q][tl])));
          const u128 nrun = add128(run, tsum);
          const double c = __ddiv_rn(dbl128(run, qb), total), ce = __ddiv_rn(dbl128(nrun, qb), total);
          const long long cb = __double_as_longlong(c), eb = __double_as_longlong(ce);
          const bool same_binade = (cb >> 52) == (eb >> 52) && ((cb >> 52) & 0x7ff) > 54;
          const double half_ulp = same_binade ? __longlong_as_double((((cb >> 52) & 0x7ff) - 53) << 52) : 0.0;
          if (same_binade && __ddiv_rn(mxw, total) < half_ulp) { run = nrun; ++tp; } else break;
        }
        s_tp = (int)tp;
        s_run = run;
        s_brk = TILE;
        if (tp >= gnt) s_jfinal = Ng - 1;
      }
      __syncthreads();
      if (s_jfinal >= 0) break;
      const int tp = s_tp;
      const int q = tp / nt, tl = tp % nt;
      const double c = dbl128(s_run, qb);
      WeightSrc<real> wq{(q == RNK) ? logw : reinterpret_cast<const real*>(pr.logw[q]), direct, ps.gmax};
      tile_cdf<real, ITEMS>(wq, qb, s_run, (long long)tl * TILE, N, Ps, Ws, s_warp);
      const int tn = (int)min((long long)TILE, N - (long long)tl * TILE);
      // first element of tile tp that does NOT vanish against its predecessor's value
      for (int j = threadIdx.x; j < tn; j += TILE_THREADS) {
        double prev = (j == 0) ? c : Ps[phys<ITEMS>(j - 1)];
        if (!vanishes(prev, Ws[phys<ITEMS>(j)], total)) atomicMin(&s_brk, j);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        if (s_brk < tn) s_jfinal = (long long)q * N + (long long)tl * TILE + s_brk - 1;
        else if (tp == gnt - 1) s_jfinal = Ng - 1;
        else {
          const u128 tsum = ld_gpu128((q == RNK) ? &tb.tile_sum[tl] : &pr.tile_sum[q][tl]);  // L2: another block wrote it
          s_run = add128(s_run, tsum);
          s_tp = tp + 1;
        }
      }
      __syncthreads();
      if (s_jfinal >= 0) break;
    }
    const long long jfinal = s_jfinal;
    for (long long i = pend + threadIdx.x; i < hi; i += TILE_THREADS) {
      if (RK > 1) {
        const unsigned q = owner_of(pr, (unsigned)i);
        pr.anc[q][i - (long long)q * N] = (int32_t)jfinal;
        wrote_remote |= (q != RNK);
      } else {
        pr.anc[0][i] = (int32_t)jfinal;
      }
    }
  }
  }  // usable
  return wrote_remote;
}

// ---------------------------------------------------------------------------------------------
// Warp-synchronous scan + search.  A thread owns ITEMS consecutive particles, a warp 32*ITEMS, and
// everything a particle needs -- its exact cumulative value, its offspring count, the end of the run
// of repeated keys it starts -- lives in registers.  One block barrier publishes the warp totals;
// from there on every warp expands ITS OWN particles into ITS OWN output range (the range is known
// from the count at the warp's first cumulative value, no exchange of counts) through a private
// shared-memory window: heads are scattered, each row of 32 outputs is filled by one ballot + one
// indexed shuffle, and the row leaves as one coalesced store.  No cumulative values or weights in
// shared memory (10 KB per block instead of 48), no block barrier in the expansion.
// cdf_out == NULL: search (systematic / stratified), writes ancestors;
// cdf_out != NULL: write the un-normalised CDF (multinomial), no search
// ---------------------------------------------------------------------------------------------
template <int ITEMS>
struct K3Smem {
  static constexpr int TILE = TILE_THREADS * ITEMS;
  static constexpr int NW = TILE_THREADS / 32;
  static constexpr int ROWS = ITEMS + (ITEMS >= 8 ? 2 : 1);  // rows of 32 outputs a warp stages per pass: its 32*ITEMS
  static constexpr int WINW = 32 * ROWS;                      // particles have about 32*ITEMS offspring
  int32_t s_res[NW][WINW];
  u128 s_warp[NW];
  u128 s_excl, s_tot, s_q, s_run;
  unsigned long long s_key;
  long long s_pend, s_jfinal;
  double s_wnext, s_u;
  int s_tp, s_brk;
  unsigned s_minw[NW];  // per-warp min weight of the tile (fp32 bits; non-negative floats order as integers)
  int s_wbrk[NW];       // first particle of the warp that starts a new key (tiles with vanishing weights only)
  static constexpr int HQ = 16;  // heavy particles (>= HEAVY offspring) of the tile: their ranges are filled by the whole block
  int s_hq[HQ][3];      // first output, end, value
  int s_hn;
};

// One tile in registers.  The scan has a LOCAL part that needs nothing but the weights -- the thread's running exact sums,
// the warp-inclusive scan of the thread totals, the warp totals in shared memory, the warp's smallest weight -- and a
// FINISH that needs the exact sum of everything before the tile: the cumulative values P_j = dbl128(exact prefix) of the
// thread's ITEMS consecutive particles, the value before the thread's / the warp's first particle, and the exact sum at
// the end of the tile.  Between the two the block must synchronise once (s_warp, s_minw).  The single-launch series
// kernel runs the local part BEFORE the grid-wide exchange of the tile sums (the tile sum is its by-product) and only
// the finish after it.
template <typename real, int ITEMS>
struct TileScan {
  typename WeightSrc<real>::wt w[ITEMS];
  u128 e[ITEMS];  // inclusive exact sums within the thread
  u128 incl;      // inclusive scan of the thread totals within the warp
};
template <typename real, int ITEMS>
struct TileRegs {
  double P[ITEMS];
  double P_tstart, P_wstart;
  u128 tile_end;
};
// weights in sc.w; base = global index of the thread's first particle; all threads call.  No barrier inside.
template <typename real, int ITEMS>
__device__ __forceinline__ void tile_scan_local(int qb, long long base, long long N, u128* s_warp, unsigned* s_minw,
                                                TileScan<real, ITEMS>& sc) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (s_minw != nullptr) {  // a lower bound (fp32, rounded down) of the warp's smallest weight
    float m = 3.4028234663852886e38f;
    if (base + ITEMS <= N) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) m = fminf(m, to_float_rd(sc.w[j]));
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j)
        if (base + j < N) m = fminf(m, to_float_rd(sc.w[j]));
    }
    const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0) s_minw[wid] = mb;
  }
  u128 run = make_u128(0, 0);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    run = add128(run, WeightSrc<real>::fix(sc.w[j], qb));
    sc.e[j] = run;
  }
  u128 incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u128 o = shfl_up128(incl, d);
    if (lane >= d) incl = add128(incl, o);
  }
  sc.incl = incl;
  if (lane == 31) s_warp[wid] = incl;
}
// after the barrier that follows tile_scan_local
template <typename real, int ITEMS>
__device__ __forceinline__ void tile_scan_finish(int qb, u128 excl, const u128* s_warp, const TileScan<real, ITEMS>& sc,
                                                 TileRegs<real, ITEMS>& r, const u128* s_woff = nullptr) {
  constexpr int NW = TILE_THREADS / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u128 woff = excl, tend = excl;
  if (s_woff != nullptr) {  // the caller already holds the exclusive prefix of the warp totals ([NW]: their sum)
    woff = add128(excl, s_woff[wid]);
    tend = add128(excl, s_woff[NW]);
  } else {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const u128 v = s_warp[w];
      tend = add128(tend, v);
      woff = add128(woff, (w < wid) ? v : make_u128(0, 0));
    }
  }
  u128 ex = shfl_up128(sc.incl, 1);
  if (lane == 0) ex = make_u128(0, 0);
  const u128 off = add128(woff, ex);
  r.P_wstart = dbl128(woff, qb);
  r.P_tstart = dbl128(off, qb);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) r.P[j] = dbl128(add128(off, sc.e[j]), qb);
  r.tile_end = tend;
}

// ONE thread, before the barrier that precedes the search: the resampling uniform (model/Resampling.scala:66) and the
// per-tile bookkeeping
template <int ITEMS>
__device__ __forceinline__ void k3_prepare(K3Smem<ITEMS>& sm, const FilterScalars* sc, const K3Ctl& ctl) {
  double u;
  if (ctl.use_u_inj) {
    u = sc->u_inj;
  } else {
    uint4 v = philox4x32(make_uint4(0u, 0u, ctl.step, RNG_RESAMPLE), ctl.key0, ctl.key1);
    u = u64_to_unit_double(v.x, v.y);
  }
  sm.s_u = u;
  sm.s_pend = 0x7FFFFFFFFFFFFFFFll;
  sm.s_hn = 0;
}

// Tile t of the scan + search once the exact sums are known: `tot` / `qsum` = sum of fix(w1) / of
// fix(w1^2) over the whole filter, `key` = ordered key of the max log-weight, `excl` = exact sum of
// everything before the tile.  Block 0 also updates ll and ESS.  PROTO3: called from the
// three-launch step (zeroes the accumulators of the next observed step); the single-launch series
// kernel keeps its own.  All threads of the block call; returns whether a peer's memory was written.
// `pre` != NULL: the caller has already run tile_scan_local for this tile (the series kernel does it before the exchange of
// the tile sums), called k3_prepare and stored sm.s_wnext, all followed by a block barrier: no barrier is needed here.
template <typename real, int ITEMS, int KIND, bool PROTO3>
__device__ __forceinline__ bool k3_tile(K3Smem<ITEMS>& sm, const real* __restrict__ logw, const double* __restrict__ direct,
                                        long long N, FilterScalars* __restrict__ sc, const SumTables& tb, const Peers& pr,
                                        const K3Ctl& ctl, const double* __restrict__ uarr, double* __restrict__ cdf_out, int t,
                                        u128 tot, u128 qsum, unsigned long long key, u128 excl,
                                        const TileScan<real, ITEMS>* pre = nullptr, const u128* pre_woff = nullptr) {
  constexpr int TILE = TILE_THREADS * ITEMS;
  constexpr int NW = K3Smem<ITEMS>::NW;
  constexpr int ROWS = K3Smem<ITEMS>::ROWS;
  constexpr int WINW = K3Smem<ITEMS>::WINW;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nt = tb.nt;
  const long long Ng = (long long)pr.R * N;  // outputs of the whole (possibly sharded) filter
  const PreScan ps = pre_scan(key, direct != nullptr);
  const int qb = ps.qb;
  const double total = dbl128(tot, qb);

  // ---- thread 0: the resampling uniform; block 0: ll increment max + log(mean w1), ESS = floor(1/sum wn^2); zero the
  //      accumulators of the next observed step.  Read after the barrier of the scan (pre: the caller did it). ----
  if (threadIdx.x == 0 && pre == nullptr) k3_prepare<ITEMS>(sm, sc, ctl);
  if (threadIdx.x == 64 && t == 0 && !ctl.defer_ll)  // another warp than the one deriving the uniform: the two run side by side
    ll_ess_update<real, PROTO3>(sc, ctl, tot, qsum, key, Ng, direct != nullptr);
  if (PROTO3 && t < tb.ns && threadIdx.x == 1) {
    tb.super_sum[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
    tb.super_q[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
  }

  WeightSrc<real> ws{logw, direct, ps.gmax};
  const long long tile0 = (long long)t * TILE;
  const int tile_n = (int)min((long long)TILE, N - tile0);
  const bool last_tile = (t == nt - 1) && (pr.rank == pr.R - 1);
  if (threadIdx.x == 32 && cdf_out == nullptr && pre == nullptr) {
    // first weight after this tile (next tile, possibly the next rank's first particle); loaded here so
    // that its latency hides behind the tile scan
    double wn = 0.0;
    if (t < nt - 1) wn = (double)ws(tile0 + TILE);
    else if (pr.rank < pr.R - 1) wn = (double)WeightSrc<real>{reinterpret_cast<const real*>(pr.logw[pr.rank + 1]), nullptr, ps.gmax}(0);
    sm.s_wnext = wn;
  }
  TileRegs<real, ITEMS> r;
  TileScan<real, ITEMS> own;
  const TileScan<real, ITEMS>* scan = pre;
  if (pre == nullptr) {
    const long long base = tile0 + (long long)threadIdx.x * ITEMS;
    ws.template load<ITEMS>(base, 1, N, own.w);
    tile_scan_local<real, ITEMS>(qb, base, N, sm.s_warp, sm.s_minw, own);
    __syncthreads();
    scan = &own;
  }
  K3_STAMP(8)
  tile_scan_finish<real, ITEMS>(qb, excl, sm.s_warp, *scan, r, pre != nullptr ? pre_woff : nullptr);
  K3_STAMP(9)

  if (cdf_out != nullptr) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int idx = threadIdx.x * ITEMS + j;
      if (idx < tile_n) cdf_out[tile0 + idx] = r.P[j];
    }
    return false;
  }

  bool wrote_remote = false;
  // all weights zero / NaN (the reference divides by a zero total here and fails later): keep every
  // particle as its own ancestor; FLAG_ZERO_TOTAL is already raised
  const bool usable = (total > 0.0) && (total - total == 0.0);
  unsigned long long* const anc64 = PROTO3 ? nullptr : ctl.anc64;  // tagged ancestors: single-rank series kernel only
  const unsigned long long tagw = (unsigned long long)ctl.anc_tag << 32;
  if (!usable) {
    for (int j = threadIdx.x; j < tile_n; j += TILE_THREADS) {
      if (anc64 != nullptr) st_relaxed_gpu(&anc64[tile0 + j], tagw | (unsigned long long)(unsigned)(tile0 + j));
      else pr.anc[pr.rank][tile0 + j] = (int32_t)((long long)pr.rank * N + tile0 + j);
    }
    return false;
  }
  // a tile whose weights are all zero owns no output: every cumulative value equals the one before the tile (a run that
  // passes through it is resolved by the tile that holds its head).  Degenerate clouds consist mostly of such tiles.
  if (!last_tile && r.tile_end.lo == excl.lo && r.tile_end.hi == excl.hi) return false;  // block-uniform
  KFun<KIND> kf{sm.s_u, (double)Ng, ctl.inv_n, total, uarr, ctl.key0, ctl.key1, ctl.step};
  const double c_end = dbl128(r.tile_end, qb);  // cumulative value of the tile's last particle (padding weighs nothing)
  const long long gbase = (long long)pr.rank * N + tile0;  // global index of the tile's first particle
  // ---- offspring counts: c_j = #{outputs with key <= P_j}; particle j owns the outputs [c_{j-1}, c_j) ----
  const double scale = __ddiv_rn((double)Ng, total);  // every thread: one division instead of a broadcast through shared memory
  const long long lo = (t == 0 && pr.rank == 0) ? 0 : kf.count_fast(dbl128(excl, qb), scale, Ng);
  const double lo_d = (double)lo;
  const int n_rel = (int)(Ng - lo);
  int cr[ITEMS];  // counts relative to lo
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int idx = threadIdx.x * ITEMS + j;
    cr[j] = (last_tile && idx >= tile_n - 1) ? n_rel : kf.count_rel(r.P[j], scale, Ng, lo, lo_d, n_rel);
  }
  // the count before the thread's first particle: the previous thread's last one; the warp's first thread evaluates
  // the count at the warp's first cumulative value itself (the same integer the previous warp's last thread converts)
  int cprev = __shfl_up_sync(FULL, cr[ITEMS - 1], 1);
  {
    const int idx_before = wid * 32 * ITEMS - 1;
    const int c_w = (wid == 0) ? 0 : ((last_tile && idx_before >= tile_n - 1) ? n_rel : kf.count_rel(r.P_wstart, scale, Ng, lo, lo_d, n_rel));
    if (lane == 0) cprev = c_w;
  }
  const int c0 = __shfl_sync(FULL, cprev, 0), c1 = __shfl_sync(FULL, cr[ITEMS - 1], 31);  // the warp's outputs [c0, c1)
  const int n_out = last_tile ? n_rel : kf.count_rel(c_end, scale, Ng, lo, lo_d, n_rel);
  const long long hi = lo + n_out;
  K3_STAMP(10)
  // A key can only repeat where a weight is at most 2^-52 of the cumulative value before it (vanishes()).  If even the
  // smallest weight of the tile is above 2^-52 of the tile's LAST cumulative value, no key of this tile repeats.
  unsigned mm = sm.s_minw[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) mm = min(mm, sm.s_minw[w]);
  const bool novanish = ctl.tie_first || (double)__uint_as_float(mm) > c_end * 2.220446049250313e-16;
  // does the run of repeated keys at the end of this tile continue into the next tile?
  const bool cont = !ctl.tie_first && !last_tile && vanishes(c_end, sm.s_wnext, total);
  if (last_tile && threadIdx.x == 0 && kf(Ng - 1) > c_end) atomicOr(&sc->flags, FLAG_CLAMPED);  // reference would throw (m.head)

  // ---- TreeMap: a duplicated key keeps the last particle inserted.  Particle p "breaks" when it starts a new key
  //      (its weight does not vanish against the cumulative value before it); the offspring of j go to the particle
  //      before the next break after j.  bmask: breaks among the thread's particles; nb_after: first break behind them.
  unsigned bmask = (1u << ITEMS) - 1u;
  int nb_after = 0;
  if (!novanish) {  // block-uniform
    bmask = 0u;
    int first_brk = 0x7FFFFFFF;
#pragma unroll
    for (int j = ITEMS - 1; j >= 0; --j) {
      const int idx = threadIdx.x * ITEMS + j;
      const double before = j ? r.P[j - 1] : r.P_tstart;
      const bool brk = (idx < tile_n) && !vanishes(before, (double)scan->w[j], total);
      if (brk) {
        bmask |= 1u << j;
        first_brk = idx;
      }
    }
    int sfx = first_brk;  // first break in this thread or a later one of the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_down_sync(FULL, sfx, d);
      if (lane + d < 32) sfx = min(sfx, o);
    }
    nb_after = __shfl_down_sync(FULL, sfx, 1);
    if (lane == 31) nb_after = 0x7FFFFFFF;
    if (lane == 0) sm.s_wbrk[wid] = sfx;
    __syncthreads();
    if (nb_after == 0x7FFFFFFF) {
#pragma unroll
      for (int w = 1; w < NW; ++w)
        if (w > wid) nb_after = min(nb_after, sm.s_wbrk[w]);
      if (nb_after == 0x7FFFFFFF) nb_after = tile_n;  // the run reaches the end of the tile
    }
  }
  if (cont) {  // block-uniform: the first output whose run reaches the end of the tile
    int nb = novanish ? 0 : nb_after;
#pragma unroll
    for (int j = ITEMS - 1; j >= 0; --j) {
      const int idx = threadIdx.x * ITEMS + j;
      const int prev = j ? cr[j - 1] : cprev;
      const int val = novanish ? idx : nb - 1;
      if (cr[j] > prev && val == tile_n - 1) atomicMin(&sm.s_pend, lo + prev);
      if ((bmask >> j) & 1u) nb = idx;
    }
    if (anc64 != nullptr) __syncthreads();  // s_pend is final before any tagged word leaves (below)
  }
  // tagged ancestors are final the moment they are stored: the outputs of a run that continues past the tile -- they are
  // rewritten after the walk below -- are held back by the expansion
  const int o_pend = (anc64 != nullptr && cont && sm.s_pend - lo < (long long)n_out) ? (int)(sm.s_pend - lo) : 0x7FFFFFFF;

  // ---- expansion: the warp's particles into the warp's outputs [c0, c1), WINW per pass.  Every particle with
  //      offspring drops its value at the head of its range; a row of 32 outputs takes, per lane, the nearest head at
  //      or below it (ballot + indexed shuffle), else the value carried over from the row before.  A particle with many
  //      offspring is not staged at all: its range is one value, written 32 outputs per store by the warp, or -- from
  //      HEAVY offspring on -- queued for the whole block. -----------------------------------------------------------
  K3_STAMP(11)
  constexpr int LONG_RUN = 64, HEAVY = 1024;
  auto store_out = [&](int o, int val) {
    const int32_t v = (int32_t)(gbase + val);
    if (anc64 != nullptr) {
      if (o < o_pend) st_relaxed_gpu(&anc64[lo + o], tagw | (unsigned long long)(unsigned)v);
    } else if (pr.R > 1) {  // offspring slot i belongs to rank i / N: scatter over NVLink
      const long long i = lo + o;
      const unsigned q = owner_of(pr, (unsigned)i);
      pr.anc[q][i - (long long)q * N] = v;
      wrote_remote |= (q != pr.rank);
    } else {
      (pr.anc[0] + lo)[o] = v;
    }
  };
  {
    int32_t* const res = sm.s_res[wid];
    bool has_long = false;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) has_long |= (cr[j] - (j ? cr[j - 1] : cprev)) >= LONG_RUN;
    has_long = __any_sync(FULL, has_long);
    int carry = 0;
    for (int w0 = c0; w0 < c1;) {  // warp-uniform
      if (has_long) {
        // the particle that owns output w0 (exactly one in the warp): does its range run on for LONG_RUN outputs?
        int own_end = 0, own_val = 0;
        bool mine = false;
        int nb = nb_after;
#pragma unroll
        for (int j = ITEMS - 1; j >= 0; --j) {
          const int idx = threadIdx.x * ITEMS + j;
          const int prev = j ? cr[j - 1] : cprev;
          if (prev <= w0 && w0 < cr[j]) {
            mine = true;
            own_end = cr[j];
            own_val = novanish ? idx : nb - 1;
          }
          if ((bmask >> j) & 1u) nb = idx;
        }
        const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
        own_end = __shfl_sync(FULL, own_end, src & 31);
        own_val = __shfl_sync(FULL, own_val, src & 31);
        if (src >= 0 && own_end - w0 >= LONG_RUN) {
          bool queued = false;
          if (own_end - w0 >= HEAVY) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&sm.s_hn, 1);
            slot = __shfl_sync(FULL, slot, 0);
            if (slot < K3Smem<ITEMS>::HQ) {
              if (lane == 0) { sm.s_hq[slot][0] = w0; sm.s_hq[slot][1] = own_end; sm.s_hq[slot][2] = own_val; }
              queued = true;
            }
          }
          if (!queued)
            for (int o = w0 + lane; o < own_end; o += 32) store_out(o, own_val);
          w0 = own_end;
          continue;
        }
      }
#pragma unroll
      for (int k = 0; k < ROWS; ++k) res[k * 32 + lane] = -1;
      __syncwarp();
      {
        int nb = nb_after;
#pragma unroll
        for (int j = ITEMS - 1; j >= 0; --j) {
          const int idx = threadIdx.x * ITEMS + j;
          const int prev = j ? cr[j - 1] : cprev;
          const unsigned rel = (unsigned)(prev - w0);
          if (cr[j] > prev && rel < (unsigned)WINW) res[rel] = novanish ? idx : nb - 1;
          if ((bmask >> j) & 1u) nb = idx;
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        const int o = w0 + k * 32 + lane;
        if (w0 + k * 32 < c1) {  // warp-uniform
          const int v = res[k * 32 + lane];
          const unsigned below = __ballot_sync(FULL, v >= 0) & (0xFFFFFFFFu >> (31 - lane));
          int got = __shfl_sync(FULL, v, (31 - __clz((int)below)) & 31);
          got = below ? got : carry;
          carry = __shfl_sync(FULL, got, 31);
          if (o < c1) store_out(o, got);
        }
      }
      __syncwarp();
      w0 += WINW;
    }
  }
  K3_STAMP(12)
  __syncthreads();  // the heavy queue is complete (and s_pend final)
  K3_STAMP(13)
  {
    const int hn = min(sm.s_hn, K3Smem<ITEMS>::HQ);
    for (int h = 0; h < hn; ++h) {
      const int a0 = sm.s_hq[h][0], a1 = sm.s_hq[h][1], val = sm.s_hq[h][2];
      for (int o = a0 + threadIdx.x; o < a1; o += TILE_THREADS) store_out(o, val);
    }
  }
  if (!cont) return wrote_remote;  // block-uniform
  const long long pend = sm.s_pend;
  if (pend < hi) {
    // The selected run of repeated keys continues past this tile; its last element is the ancestor.
    // Walk forward over the GLOBAL tile sequence (rank-major): whole tiles are skipped from the
    // tables when every weight in them is strictly below half an ulp of the running (normalised)
    // value and the value stays in its binade, otherwise the tile is recomputed.
    const long long gnt = (long long)pr.R * nt;
    if (threadIdx.x == 0) { sm.s_tp = pr.rank * nt + t + 1; sm.s_run = r.tile_end; sm.s_jfinal = -1; }
    __syncthreads();
    for (;;) {
      if (threadIdx.x == 0) {
        long long tp = sm.s_tp;
        u128 run = sm.s_run;
        while (tp < gnt) {
          const int q = (int)(tp / nt), tl = (int)(tp % nt);
          const u128 tsum = walk_tile_sum(tb, pr, q, tl);  // L2: another block wrote it
          const double mxw = walk_tile_maxw(tb, pr, q, tl);
          const u128 nrun = add128(run, tsum);
          const double c = __ddiv_rn(dbl128(run, qb), total), ce = __ddiv_rn(dbl128(nrun, qb), total);
          const long long cb = __double_as_longlong(c), eb = __double_as_longlong(ce);
          const bool same_binade = (cb >> 52) == (eb >> 52) && ((cb >> 52) & 0x7ff) > 54;
          const double half_ulp = same_binade ? __longlong_as_double((((cb >> 52) & 0x7ff) - 53) << 52) : 0.0;
          if (same_binade && __ddiv_rn(mxw, total) < half_ulp) { run = nrun; ++tp; } else break;
        }
        sm.s_tp = (int)tp;
        sm.s_run = run;
        sm.s_brk = TILE;
        if (tp >= gnt) sm.s_jfinal = Ng - 1;
      }
      __syncthreads();
      if (sm.s_jfinal >= 0) break;
      const int tp = sm.s_tp;
      const int q = tp / nt, tl = tp % nt;
      const u128 run0 = sm.s_run;
      WeightSrc<real> wq{(q == pr.rank) ? logw : reinterpret_cast<const real*>(pr.logw[q]), direct, ps.gmax};
      TileRegs<real, ITEMS> rq;
      TileScan<real, ITEMS> sq;
      wq.template load<ITEMS>((long long)tl * TILE + (long long)threadIdx.x * ITEMS, 1, N, sq.w);
      __syncthreads();  // s_warp may still be read from its previous use
      tile_scan_local<real, ITEMS>(qb, 0, N, sm.s_warp, nullptr, sq);
      __syncthreads();
      tile_scan_finish<real, ITEMS>(qb, run0, sm.s_warp, sq, rq);
      const int tn = (int)min((long long)TILE, N - (long long)tl * TILE);
      // first element of tile tp that does NOT vanish against its predecessor's value
#pragma unroll
      for (int j = ITEMS - 1; j >= 0; --j) {
        const int idx = threadIdx.x * ITEMS + j;
        const double before = j ? rq.P[j - 1] : rq.P_tstart;
        if (idx < tn && !vanishes(before, (double)sq.w[j], total)) atomicMin(&sm.s_brk, idx);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        if (sm.s_brk < tn) sm.s_jfinal = (long long)q * N + (long long)tl * TILE + sm.s_brk - 1;
        else if (tp == gnt - 1) sm.s_jfinal = Ng - 1;
        else {
          sm.s_run = rq.tile_end;  // thread 0 holds the exact sum at the end of tile tp
          sm.s_tp = tp + 1;
        }
      }
      __syncthreads();
      if (sm.s_jfinal >= 0) break;
    }
    const long long jfinal = sm.s_jfinal;
    for (long long i = pend + threadIdx.x; i < hi; i += TILE_THREADS) {
      if (anc64 != nullptr) {
        st_relaxed_gpu(&anc64[i], tagw | (unsigned long long)(unsigned)jfinal);
      } else if (pr.R > 1) {
        const unsigned q = owner_of(pr, (unsigned)i);
        pr.anc[q][i - (long long)q * N] = (int32_t)jfinal;
        wrote_remote |= (q != pr.rank);
      } else {
        pr.anc[0][i] = (int32_t)jfinal;
      }
    }
  }
  return wrote_remote;
}

// ---------------------------------------------------------------------------------------------
// K3 fast path: the tile in fp64 with CERTIFIED counts, the exact path as the fallback.
//
// What makes the scan + search expensive is exactness per particle: a 128-bit fixed-point prefix, its conversion to
// fp64, an offspring count that must be THE count of keys below that value.  But the count is floor(y) + 1 with
// y = P_j n / total - u, and floor() is insensitive to errors in y unless y is near an integer.  So the tile is
// scanned in plain fp64 -- prefix L_j of the weights inside the tile (relative error below 2.3e-13 after 2048 adds) on
// top of the EXACT cumulative value before the tile -- and y'_j = fma(L_j, s, fma(P_b, s, -(u + lo))) differs from the
// exact path's y_j by less than 1e-6 for every cloud the library accepts (n < 2^31).  The exact path takes the
// arithmetic count when y_j is further than 1e-5 from an integer (KFun::count_rel, proven there); the fast path takes
// it when y'_j is further than 1.2e-5 away -- then y_j is outside its own band, on the same side of the same integer,
// and both paths return the same count.  Likewise the TreeMap duplicate-key test vanishes(P, w): decided from P' when
// w is clearly above 2^-52 P' or clearly below 2^-55 P' (a relative margin of 1e-9 absorbs the 2.3e-13), undecided in
// between.  One undecided particle (4 % of the tiles at 2^24) sends the WHOLE tile to the exact path, which recomputes
// it from scratch; so do the last tile of the cloud (clamping rule, ragged end) and a tile whose trailing run of
// repeated keys continues into the next tile.  Ancestors are therefore bit-identical to the exact path, always.
// Applies to fp32 filters, 2048-particle tiles, systematic resampling, one rank -- the large-cloud configuration.
// ---------------------------------------------------------------------------------------------
// vanishes(P, w, total) decided from an APPROXIMATION Pa of P (relative error below 1e-12): 1 repeats, 0 does not,
// -1 cannot be certified.  The exact test is fl(c + wn) == c with c = fl(P / total), wn = fl(w / total): true iff wn is
// below half an ulp of c (or equal to it with an even c).  c and ca = fl(Pa / total) have the same ulp unless c sits
// within 1e-11 of a power of two; wn is the same number on both paths; so wn at least 1e-9 away (relatively) from half
// an ulp of ca settles the matter.  Only weights between 2^-55 P and 2^-52 P come here.
static __device__ __noinline__ int vanishes_certified(double Pa, double w, double total) {
  const double ca = __ddiv_rn(Pa, total), wn = __ddiv_rn(w, total);
  const long long cb = __double_as_longlong(ca);
  const int e = (int)((cb >> 52) & 0x7ff);
  const unsigned long long mant = (unsigned long long)cb & 0xFFFFFFFFFFFFFull;
  if (e < 60 || e > 2000) return -1;
  if (mant < (1ull << 14) || mant > 0xFFFFFFFFFFFFFull - (1ull << 14)) return -1;  // within 4e-12 of a binade border
  const double half_ulp = __longlong_as_double((long long)(e - 53) << 52);
  if (wn <= half_ulp * (1.0 - 1e-9)) return 1;
  if (wn >= half_ulp * (1.0 + 1e-9)) return 0;
  return -1;
}

struct K3FastSmem {
  static constexpr int TILE = TILE_THREADS * 8;
  static constexpr int WIN = TILE + TILE_THREADS;
  int32_t s_res[WIN];
  double s_wsum[TILE_THREADS / 32];
  int s_cnt[TILE_THREADS], s_cnt2[TILE_THREADS / 32];
  unsigned s_vbits[TILE / 32];  // bit p: particle p of the tile repeats the key of its predecessor
  double s_u, s_scale;
  long long s_lo;
  int s_flag, s_anyv;
};

template <bool FLAT, bool SHD = false>
__device__ __forceinline__ bool k3_tile_fast(K3FastSmem& fs, const float* __restrict__ logw, long long N, FilterScalars* __restrict__ sc,
                                             const SumTables& tb, const K3Ctl& ctl, int32_t* __restrict__ anc_out, int t, u128 tot,
                                             u128 qsum, unsigned long long key, u128 excl, int R = 1, int rank = 0) {
  // R > 1: rank `rank` of a sharded filter.  `excl` includes the ranks before, counts and ancestors are global, anc_out is
  // this rank's buffer: a tile whose outputs all fall into this rank's own slots is settled here, any other -- and the
  // last tile of every rank -- by the exact path, which scatters to the peers.
  constexpr int ITEMS = 8, TILE = K3FastSmem::TILE, WIN = K3FastSmem::WIN, PER = WIN / TILE_THREADS, NW = TILE_THREADS / 32;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nt = tb.nt;
  if (t >= nt - 1) return false;  // the last tile: clamping rule and ragged end are the exact path's
  const PreScan ps = pre_scan(key, false);
  const float gmaxf = (float)ps.gmax;
  const double total = dbl128(tot, 96);
  if (!((total > 0.0) && (total - total == 0.0))) return false;
  const long long tile0 = (long long)t * TILE;
  const double P_b = dbl128(excl, 96);
  const double c_end = dbl128(add128(excl, tb.tile_sum[t]), 96);  // the exact cumulative value of the tile's last particle
  if (!ctl.tie_first) {  // a run of repeated keys that leaves the tile is resolved by the exact path's walk
    const float w_next = expf_det(__fsub_rn(__ldg(logw + tile0 + TILE), gmaxf));
    if (vanishes(c_end, (double)w_next, total)) return false;
  }
  const long long Ng = SHD ? (long long)R * N : N, own0 = SHD ? (long long)rank * N : 0;
  if (threadIdx.x == 0) {  // one thread: the resampling uniform, n / total and the number of outputs before the tile
    double u;
    if (ctl.use_u_inj) {
      u = sc->u_inj;
    } else {
      uint4 v = philox4x32(make_uint4(0u, 0u, ctl.step, RNG_RESAMPLE), ctl.key0, ctl.key1);
      u = u64_to_unit_double(v.x, v.y);
    }
    const double scale = __ddiv_rn((double)Ng, total);
    KFun<CSSM_RESAMPLE_SYSTEMATIC> kf{u, (double)Ng, ctl.inv_n, total, nullptr, ctl.key0, ctl.key1, ctl.step};
    fs.s_u = u;
    fs.s_scale = scale;
    fs.s_lo = (t == 0 && (!SHD || rank == 0)) ? 0 : kf.count_fast(P_b, scale, Ng);
    fs.s_flag = 0;
    fs.s_anyv = 0;
  }
  if (threadIdx.x < TILE / 32) fs.s_vbits[threadIdx.x] = 0u;

  // ---- weights and their fp64 prefix inside the tile (fixed association: the same bits for every launch) ----
  float w[ITEMS];
  double L[ITEMS];
  {
    const float4 a = *reinterpret_cast<const float4*>(logw + tile0 + threadIdx.x * ITEMS);
    const float4 b = *reinterpret_cast<const float4*>(logw + tile0 + threadIdx.x * ITEMS + 4);
    const float lw[ITEMS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      w[j] = expf_det(__fsub_rn(lw[j], gmaxf));
      run = __dadd_rn(run, (double)w[j]);
      L[j] = run;
    }
  }
  double incl = L[ITEMS - 1];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl = __dadd_rn(incl, o);
  }
  if (lane == 31) fs.s_wsum[wid] = incl;
  __syncthreads();
  double off = 0.0;
#pragma unroll
  for (int ww = 0; ww < NW - 1; ++ww)
    if (ww < wid) off = __dadd_rn(off, fs.s_wsum[ww]);
  {
    const double ex = __shfl_up_sync(FULL, incl, 1);
    if (lane > 0) off = __dadd_rn(off, ex);
  }

  // ---- offspring counts, certified ----
  const double scale = fs.s_scale;
  const long long lo = fs.s_lo;
  const int n_rel = (int)(Ng - lo);
  const double yb = __fma_rn(P_b, scale, -__dadd_rn(fs.s_u, (double)lo));
  bool undecided = false;
  unsigned vmask = 0u;
  int cr[ITEMS];
  // no weight of the thread can repeat a key if even the smallest is clearly above 2^-52 of the LARGEST cumulative value
  // the thread sees: one test for eight particles in the common case
  float wmin = w[0];
#pragma unroll
  for (int j = 1; j < ITEMS; ++j) wmin = fminf(wmin, w[j]);
  const bool check_v = !ctl.tie_first && !((double)wmin > __dadd_rn(P_b, __dadd_rn(off, L[ITEMS - 1])) * 2.2204460514709114e-16);
  double cum_prev = off;  // tile-local cumulative value of the previous particle
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const double cum = __dadd_rn(off, L[j]);
    const double y = __fma_rn(cum, scale, yb);
    const int a = __double2int_rd(__dadd_rn(y, -1.2e-5)), b = __double2int_rd(__dadd_rn(y, 1.2e-5));
    undecided |= (a != b);
    cr[j] = max(min(a, n_rel - 1) + 1, 0);
    if (check_v && (threadIdx.x | j) != 0) {
      const double wd = (double)w[j], before = __dadd_rn(P_b, cum_prev);
      if (!(wd > before * 2.2204460514709114e-16)) {          // not clearly above 2^-52 P
        if (wd <= before * 2.7755575587873340e-17) vmask |= 1u << j;  // clearly below 2^-55 P: the key repeats
        else {
          const int r = vanishes_certified(before, wd, total);
          if (r > 0) vmask |= 1u << j;
          undecided |= (r < 0);
        }
      }
    }
    cum_prev = cum;
  }
  if (undecided) fs.s_flag = 1;
  if (vmask) {
    atomicOr(&fs.s_vbits[threadIdx.x >> 2], vmask << ((threadIdx.x & 3) * 8));
    fs.s_anyv = 1;
  }
  fs.s_cnt[threadIdx.x] = cr[ITEMS - 1];
  __syncthreads();
  // block-uniform; nothing has been written to global memory yet.  Sharded: outputs in a peer's slots are the exact path's.
  if (fs.s_flag || (SHD && (lo < own0 || lo + fs.s_cnt[TILE_THREADS - 1] > own0 + N))) {
    if (threadIdx.x == 0) atomicAdd(&sc->n_exact, 1ull);
    return false;
  }
  if (threadIdx.x == 0) atomicAdd(&sc->n_fast, 1ull);

  // ---- expansion, WIN outputs per pass (head scatter + max-scan, as the exact path), coalesced stores ----
  const int n_out = fs.s_cnt[TILE_THREADS - 1];
  const bool anyv = fs.s_anyv != 0;
  const int32_t gbase = SHD ? (int32_t)(own0 + tile0) : (int32_t)tile0;
  const int prev0 = threadIdx.x ? fs.s_cnt[threadIdx.x - 1] : 0;
  int carry = 0;
  constexpr int HEAVY_SPAN = 8 * WIN;  // as in the exact path: a stretch this long without a head is filled directly
  for (int w0 = 0; w0 < n_out; w0 += WIN) {
    if (w0 > 0 && n_out - w0 >= HEAVY_SPAN) {  // block-uniform, and false for every tile of an ordinary cloud
      int nh = 0x7FFFFFFF;  // the first head at or behind w0: up to there every output belongs to `carry`
      {
        int prev = prev0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          if (cr[j] > prev && prev >= w0) nh = min(nh, prev);
          prev = cr[j];
        }
      }
      nh = __reduce_min_sync(FULL, nh);
      if (lane == 0) fs.s_cnt2[wid] = nh;
      __syncthreads();
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) nh = min(nh, fs.s_cnt2[ww]);
      __syncthreads();  // s_cnt2 is used again by the pass below
      const int span_end = min(nh, n_out);
      if (span_end - w0 >= HEAVY_SPAN) {
        int jt = carry;
        if (anyv)
          while (jt + 1 < TILE && ((fs.s_vbits[(jt + 1) >> 5] >> ((jt + 1) & 31)) & 1u)) ++jt;
        const int32_t val = gbase + jt;
        int32_t* const out = anc_out + (lo - own0);
        for (int o = w0 + (int)threadIdx.x; o < span_end; o += TILE_THREADS) out[o] = val;
        w0 = span_end - WIN;  // the loop adds WIN: the next pass starts at the next particle's first output
        continue;
      }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) fs.s_res[threadIdx.x * PER + k] = -1;
    __syncthreads();
    {
      int prev = prev0;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        if (cr[j] > prev && (unsigned)(prev - w0) < (unsigned)WIN) {
          int jt = threadIdx.x * ITEMS + j;
          if (anyv)  // TreeMap: the offspring go to the last particle of the run of repeated keys that follows (rare)
            while (jt + 1 < TILE && ((fs.s_vbits[(jt + 1) >> 5] >> ((jt + 1) & 31)) & 1u)) ++jt;
          fs.s_res[prev - w0] = jt;
        }
        prev = cr[j];
      }
    }
    __syncthreads();
    int v[PER];
    int run = -1;
#pragma unroll
    for (int k = 0; k < PER; ++k) { run = max(run, fs.s_res[threadIdx.x * PER + k]); v[k] = run; }
    int inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc = max(inc, o); }
    if (lane == 31) fs.s_cnt2[wid] = inc;
    __syncthreads();
    int before = carry;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      const int c2 = fs.s_cnt2[ww];
      if (ww < wid) before = max(before, c2);
      carry = max(carry, c2);
    }
    { const int o = __shfl_up_sync(FULL, inc, 1); if (lane > 0) before = max(before, o); }
#pragma unroll
    for (int k = 0; k < PER; ++k) fs.s_res[threadIdx.x * PER + k] = max(v[k], before);
    __syncthreads();
    const int n_w = min(WIN, n_out - w0);
    int32_t* const out = SHD ? anc_out + (lo - own0 + w0) : anc_out + (lo + w0);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int o = threadIdx.x + k * TILE_THREADS;
      if (o < n_w) out[o] = gbase + fs.s_res[o];
    }
    __syncthreads();
  }
  // ---- what block 0 / the first blocks do besides their tile (the exact path does the same at its top) ----
  if (threadIdx.x == 64 && t == 0) ll_ess_update<float, true>(sc, ctl, tot, qsum, key, Ng, false);
  if (!FLAT && t < tb.ns && threadIdx.x == 1) {
    tb.super_sum[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
    tb.super_q[(size_t)(ctl.parity ^ 1) * tb.ns + t] = make_u128(0, 0);
  }
  return true;
}

#ifndef CSSM_K3_MINBLOCKS
#ifdef CSSM_K3_WS
#define CSSM_K3_MINBLOCKS 3
#else
#define CSSM_K3_MINBLOCKS 4
#endif
#endif
// FLAT: the sum tables have no super tiles (SumTables::ns == 0, single rank) -- a separate instantiation, so that the
// two-level kernel of the large clouds keeps its register allocation
template <typename real, int ITEMS, int KIND, bool FLAT = false, bool SH = false>
__global__ void __launch_bounds__(TILE_THREADS, CSSM_K3_MINBLOCKS)
k_scan_search(const real* __restrict__ logw, const double* __restrict__ direct, long long N, FilterScalars* __restrict__ sc,
              SumTables tb, const __grid_constant__ Peers pr, K3Ctl ctl, const double* __restrict__ uarr,
              double* __restrict__ cdf_out, int pub_here = 0) {
#ifdef CSSM_K3_WS
  __shared__ K3Smem<ITEMS> sm;
#else
  // the fast path's window and the exact path's tile never live at the same time; the few words both need
  // (s_excl, s_tot, s_q, s_key) are kept outside the union
  constexpr bool FAST = std::is_same<real, float>::value && ITEMS == 8 && KIND == CSSM_RESAMPLE_SYSTEMATIC;
  __shared__ union SmemU {
    K3SmemBlk<ITEMS> blk;
    K3FastSmem fast;
  } smem_u;
  K3SmemBlk<ITEMS>& sm = smem_u.blk;
#endif
  u128& s_excl = sm.s_excl;
  u128& s_tot = sm.s_tot;
  u128& s_q = sm.s_q;
  unsigned long long& s_key = sm.s_key;
  u128* s_warp = sm.s_warp;

  griddep_wait();
  griddep_launch();
  const int t = blockIdx.x, p = ctl.parity;
  const int RK = SH ? pr.R : 1;
  StepAcc* A = &sc->acc[p];

  // ---- totals: this rank's from the accumulators, the other ranks' from the exchange slots ------
  if (RK > 1) {
    if (pub_here && blockIdx.x == 0 && threadIdx.x < 32) publish_sums_warp(pr, sc, p, ctl.obs_seq);  // K2 is complete: this rank's exact sums
    const XchSlot* mine = pr.xch[pr.rank];
    gate_wait(&sc->gate3, ctl.obs_seq + 1, pr, sc, [&](int q) { return &mine[q].sum_seq[p]; });
    if (threadIdx.x < 32) {  // lane q reads rank q's slot; sums by shuffles
      const int q = threadIdx.x;
      const bool on = q < RK;
      u128 tq = on ? make_u128(ld_relaxed_sys(&mine[q].tot_lo[p]), ld_relaxed_sys(&mine[q].tot_hi[p])) : make_u128(0, 0);
      u128 qq = on ? make_u128(ld_relaxed_sys(&mine[q].q_lo[p]), ld_relaxed_sys(&mine[q].q_hi[p])) : make_u128(0, 0);
      unsigned long long key = on ? ld_relaxed_sys(&mine[q].max_key[p]) : 0ull;
      u128 before = (on && q < pr.rank) ? tq : make_u128(0, 0);
      tq = warp_sum128(tq);
      qq = warp_sum128(qq);
      before = warp_sum128(before);
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, m);
        key = o > key ? o : key;
      }
      if (threadIdx.x == 0) { s_tot = tq; s_q = qq; s_key = key; s_excl = before; }
    }
  } else if (threadIdx.x == 0) {
    s_tot = A->tot;
    s_q = A->q;
    s_key = A->gmax_key;
    s_excl = make_u128(0, 0);
  }
  if (FLAT) {
    // flat mode (single rank): totals and the exclusive prefix straight from the tile sums
    u128 at = make_u128(0, 0), aq = make_u128(0, 0), ae = make_u128(0, 0);
    for (int tt = threadIdx.x; tt < tb.nt; tt += TILE_THREADS) {
      const u128 v = tb.tile_sum[tt];
      at = add128(at, v);
      if (tt < t) ae = add128(ae, v);
    }
    at = warp_sum128(at);
    ae = warp_sum128(ae);
    if (t == 0) {  // the sum of squares only feeds the ESS, which block 0 computes
      for (int tt = threadIdx.x; tt < tb.nt; tt += TILE_THREADS) aq = add128(aq, tb.tile_q[tt]);
      aq = warp_sum128(aq);
    }
    __shared__ u128 s_r3[3][TILE_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { s_r3[0][threadIdx.x >> 5] = at; s_r3[1][threadIdx.x >> 5] = aq; s_r3[2][threadIdx.x >> 5] = ae; }
    __syncthreads();
    if (threadIdx.x == 0) {
      u128 r0 = s_r3[0][0], r1 = s_r3[1][0], r2 = s_r3[2][0];
      for (int w = 1; w < TILE_THREADS / 32; ++w) {
        r0 = add128(r0, s_r3[0][w]);
        r1 = add128(r1, s_r3[1][w]);
        r2 = add128(r2, s_r3[2][w]);
      }
      s_tot = r0;
      s_q = r1;
      s_excl = r2;
    }
  } else
  // ---- exact sum of everything before this tile: whole super tiles + the tiles of this super ----
  {
    const int sidx = t / SUPER;
    const u128* ssum = tb.super_sum + (size_t)p * tb.ns;
    u128 acc = make_u128(0, 0);
    for (int s = threadIdx.x; s < sidx; s += TILE_THREADS) acc = add128(acc, ssum[s]);
    for (int tt = sidx * SUPER + threadIdx.x; tt < t; tt += TILE_THREADS) acc = add128(acc, tb.tile_sum[tt]);
    const u128 local = block_sum128(acc, s_warp);
    if (threadIdx.x == 0) s_excl = add128(s_excl, local);
  }
  __syncthreads();
#ifdef CSSM_K3_WS
  const bool wrote_remote = k3_tile<real, ITEMS, KIND, true>(sm, logw, direct, N, sc, tb, pr, ctl, uarr, cdf_out, t, s_tot, s_q,
                                                             s_key, s_excl);
#else
  const u128 v_tot = s_tot, v_q = s_q, v_excl = s_excl;  // to registers: the fast path reuses the shared memory they sit in
  const unsigned long long v_key = s_key;
  if (SH && RK > 1 && t == 0 && threadIdx.x == 96 && cdf_out == nullptr) {
    // Which global output slots belong to THIS rank's particles: [count of keys <= the exact sum before the rank,
    // count of keys <= the exact sum at its end) -- the same counts its first and last tile use.  The next K1 lets
    // blocks inside that range start without waiting for the peers (their ancestors are this kernel's own work).
    const PreScan ps = pre_scan(v_key, direct != nullptr);
    const double total = dbl128(v_tot, ps.qb);
    const long long Ng = (long long)RK * N;
    long long lo = (long long)pr.rank * N, hi = lo + N;  // unusable total: every particle stays its own ancestor
    if ((total > 0.0) && (total - total == 0.0)) {
      double u;
      if (ctl.use_u_inj) {
        u = sc->u_inj;
      } else {
        uint4 v = philox4x32(make_uint4(0u, 0u, ctl.step, RNG_RESAMPLE), ctl.key0, ctl.key1);
        u = u64_to_unit_double(v.x, v.y);
      }
      KFun<KIND> kf{u, (double)Ng, ctl.inv_n, total, uarr, ctl.key0, ctl.key1, ctl.step};
      const double scale = __ddiv_rn((double)Ng, total);
      lo = (pr.rank == 0) ? 0 : kf.count_fast(dbl128(v_excl, ps.qb), scale, Ng);
      hi = (pr.rank == RK - 1) ? Ng : kf.count_fast(dbl128(add128(v_excl, A->tot), ps.qb), scale, Ng);
    }
    sc->out_lo = lo;
    sc->out_hi = hi;
  }
  bool settled = false;  // sharded only: the certified path wrote the tile's ancestors (all of them in this rank's slots)
  if (FAST && direct == nullptr && cdf_out == nullptr && uarr == nullptr && ctl.fast_ok) {
    __syncthreads();  // everyone holds the four values
    if (!SH) {
      if (k3_tile_fast<FLAT>(smem_u.fast, reinterpret_cast<const float*>(logw), N, sc, tb, ctl, pr.anc[0], t, v_tot, v_q, v_key, v_excl))
        return;
    } else {
      settled = k3_tile_fast<FLAT, true>(smem_u.fast, reinterpret_cast<const float*>(logw), N, sc, tb, ctl, pr.anc[pr.rank], t, v_tot, v_q,
                                   v_key, v_excl, RK, pr.rank);
    }
    __syncthreads();  // undecided: the exact path recomputes the tile from scratch
  }
  bool wrote_remote = false;
  if (!SH || !settled)
    wrote_remote = k3_tile_blk<real, ITEMS, KIND, true, SH>(sm, logw, direct, N, sc, tb, pr, ctl, uarr, cdf_out, t, v_tot, v_q, v_key,
                                                            v_excl);
#endif
  if (cdf_out != nullptr) return;
  // Sharded: "resampling done" is told to the peers by block 0 of the next K1 (stream order: this kernel is complete by then)
  (void)wrote_remote;
}

// K4'  multinomial: Breeze Multinomial.draw first-draw walk = first j with cumulative >= u*sum
static __global__ void __launch_bounds__(256)
k_multinomial_search(const double* __restrict__ cdf, long long N, const double* __restrict__ uarr, uint32_t key0,
                     uint32_t key1, uint32_t step, int32_t* __restrict__ anc, int* __restrict__ flags_out) {
  griddep_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double u;
  if (uarr) u = uarr[i];
  else {
    uint4 v = philox4x32(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), step, RNG_RESAMPLE | 1u), key0, key1);
    u = u64_to_unit_double(v.x, v.y);
  }
  const double target = __dmul_rn(u, cdf[N - 1]);
  long long a = 0, b = N;
  while (a < b) {
    long long m = (a + b) >> 1;
    if (cdf[m] >= target) b = m; else a = m + 1;
  }
  if (a >= N) { a = 0; atomicOr(flags_out, FLAG_CLAMPED); }
  anc[i] = (int32_t)a;
}

// ---------------------------------------------------------------------------------------------
// K5  gather (only when the resampled cloud has to be materialised: get_particles, tests)
//     anc holds global indices; pr tells where each parent lives
// ---------------------------------------------------------------------------------------------
template <typename real, typename out_t>
__global__ void __launch_bounds__(256)
k_gather(const __grid_constant__ Peers pr, const int32_t* __restrict__ anc, out_t* __restrict__ out, int d, long long N,
         long long Ns, long long out_stride) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const real* src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
  if (anc) {
    const unsigned g = (unsigned)anc[i];
    const unsigned q = (pr.R > 1) ? owner_of(pr, g) : 0u;
    src = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
  }
  for (int k = 0; k < d; ++k) out[(long long)k * out_stride + i] = (out_t)src[(long long)k * Ns];
}

// copy with dtype conversion (logw / propagated state read-back)
template <typename real>
__global__ void __launch_bounds__(256) k_to_double(const real* __restrict__ in, double* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}
template <typename real>
__global__ void __launch_bounds__(256)
k_w1_out(const real* __restrict__ logw, const FilterScalars* __restrict__ sc, double* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)WeightSrc<real>{logw, nullptr, sc->gmax}(i);
}

// Resampling.sampleOne (model/Resampling.scala:151-154): one uniformly chosen particle of this
// rank's part of the current (resampled) cloud -> out[d] (double)
template <typename real>
__global__ void k_sample_one(const __grid_constant__ Peers pr, const int32_t* __restrict__ anc, double* __restrict__ out,
                             int d, long long N, long long Ns, uint32_t key0, uint32_t key1, uint32_t step) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint4 v = philox4x32(make_uint4(0u, 0u, step, RNG_SAMPLE_ONE), key0, key1);
  unsigned long long r = ((unsigned long long)v.x << 32) | v.y;
  long long i = (long long)(r % (unsigned long long)N);
  const real* src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
  if (anc) {
    const unsigned g = (unsigned)anc[i];
    const unsigned q = (pr.R > 1) ? owner_of(pr, g) : 0u;
    src = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
  }
  for (int k = 0; k < d; ++k) out[k] = (double)src[(long long)k * Ns];
}

// per-coordinate mean of the resampled cloud (ParticleFilter.meanState); fp64 accumulation
template <typename real>
__global__ void __launch_bounds__(256)
k_mean_state(const __grid_constant__ Peers pr, const int32_t* __restrict__ anc, double* __restrict__ out, int d, long long N,
             long long Ns) {
  const int k = blockIdx.y;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const real* src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
    if (anc) {
      const unsigned g = (unsigned)anc[i];
      const unsigned q = (pr.R > 1) ? owner_of(pr, g) : 0u;
      src = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
    }
    acc += (double)src[(long long)k * Ns];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + k, acc / (double)N);
}

// ---------------------------------------------------------------------------------------------
// Order statistics of the cloud without sorting it (ParticleFilter.getIntervals /
// getCredibleInterval / getOrderStatistic, model/ParticleFilter.scala:415-424,455-460,490-505):
// MSD radix select, 8 bits per pass, on the order-preserving integer image of the values.  Column
// c < d is coordinate c of the resampled cloud, column d is gamma = f(x, t) = sum_k C[k] x[k]
// (evaluated with the same fma chain as K1).  Every column carries TWO targets (lower / upper
// rank); a pass histograms, per target, the next digit of the elements that match the digits
// chosen so far, k_select_pick then fixes that digit.  The result is an element of the cloud,
// bit for bit -- selection, unlike the reference's full sort, is O(N) per pass.
// ---------------------------------------------------------------------------------------------
struct SelState {
  unsigned long long prefix;  // digits chosen so far (right-aligned)
  long long rank;             // rank of the target among the elements that match the prefix
};
template <typename real> struct KeyOf;
template <> struct KeyOf<float> {
  typedef unsigned type;
  static constexpr int BITS = 32;
  __device__ __forceinline__ static unsigned key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  __device__ __forceinline__ static double value(unsigned long long k) {
    const unsigned kk = (unsigned)k;
    return (double)__uint_as_float((kk & 0x80000000u) ? (kk & 0x7FFFFFFFu) : ~kk);
  }
};
template <> struct KeyOf<double> {
  typedef unsigned long long type;
  static constexpr int BITS = 64;
  __device__ __forceinline__ static unsigned long long key(double v) { return ord_key(v); }
  __device__ __forceinline__ static double value(unsigned long long k) { return ord_unkey(k); }
};

template <typename real>
__global__ void __launch_bounds__(256)
k_select_hist(const __grid_constant__ Peers pr, const int32_t* __restrict__ anc, const __grid_constant__ StepArgs<real> a, int d,
              long long N, long long Ns, int pass, const SelState* __restrict__ sel, unsigned* __restrict__ hist) {
  typedef typename KeyOf<real>::type key_t;
  constexpr int BITS = KeyOf<real>::BITS;
  __shared__ unsigned h[2][256];
  const int c = blockIdx.y;  // column: coordinate c, or gamma when c == d
  h[0][threadIdx.x] = 0u;
  h[1][threadIdx.x] = 0u;
  __syncthreads();
  const int shift = BITS - 8 * (pass + 1);  // position of this pass's digit
  const key_t p0 = (key_t)sel[2 * c].prefix, p1 = (key_t)sel[2 * c + 1].prefix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const real* src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
    if (anc) {
      const unsigned g = (unsigned)anc[i];
      const unsigned q = (pr.R > 1) ? owner_of(pr, g) : 0u;
      src = reinterpret_cast<const real*>(pr.x[q]) + (g - q * (unsigned)pr.Nl);
    }
    real v;
    if (c < d) {
      v = src[(long long)c * Ns];
    } else {
      v = (real)0;
      for (int k = 0; k < d; ++k) v = r_fma<real>(a.C[k], src[(long long)k * Ns], v);
    }
    const key_t key = KeyOf<real>::key(v);
    const key_t hi = (pass == 0) ? (key_t)0 : (key_t)(key >> (shift + 8));
    const unsigned digit = (unsigned)(key >> shift) & 0xFFu;
    if (hi == p0) atomicAdd(&h[0][digit], 1u);
    if (hi == p1) atomicAdd(&h[1][digit], 1u);
  }
  __syncthreads();
  unsigned* out = hist + (size_t)c * 512;
  if (h[0][threadIdx.x]) atomicAdd(out + threadIdx.x, h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(out + 256 + threadIdx.x, h[1][threadIdx.x]);
}

// one block of 32 threads per (column, target): the digit whose bin holds the target rank
static __global__ void __launch_bounds__(32) k_select_pick(SelState* __restrict__ sel, unsigned* __restrict__ hist) {
  const int ct = blockIdx.x;  // 2*c + target
  unsigned* hh = hist + (size_t)ct * 256;
  if (threadIdx.x == 0) {
    SelState st = sel[ct];
    long long cum = 0;
    int digit = 255;
    for (int b = 0; b < 256; ++b) {
      const long long n = (long long)hh[b];
      if (st.rank < cum + n) { digit = b; break; }
      cum += n;
    }
    st.prefix = (st.prefix << 8) | (unsigned long long)digit;
    st.rank -= cum;
    sel[ct] = st;
  }
  __syncwarp();
  for (int b = threadIdx.x; b < 256; b += 32) hh[b] = 0u;  // ready for the next pass
}

template <typename real>
__global__ void k_select_finish(const SelState* __restrict__ sel, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = KeyOf<real>::value(sel[i].prefix);
}

}  // namespace cssm
