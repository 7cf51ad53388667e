// cssm_series.cuh -- llFilter (model/ParticleFilter.scala:137-140) of a SMALL cloud as ONE launch: k_series_one (at most 1024
// particles, one block), k_series_small (one block per 512-particle tile, all resident; described first), k_series_multi.
//
// A PMMH likelihood evaluation (model/PMMH.scala:71) filters a few tens of thousands of particles
// over hundreds of observations: every stage of a step is a few microseconds of work and the
// three-launch step of cssm_kernels.cuh is bound by launch and dependency latency (19 us per
// observation at 2^16 particles).  Here the whole foldLeft over the observations runs inside one
// cooperative kernel: one block per 512-particle tile, all blocks resident, the three stages of a
// step separated by grid-wide barriers instead of kernel boundaries
//
//   P1  gather + stepFunction + f + dataLikelihood, block max -> atomicMax      PF :118,:123-124
//   --  barrier
//   P2  w1 = exp(logw - max), exact tile sums of w1 and w1^2                    :125, :431-434
//   --  barrier
//   P3  every block adds the (<= a few hundred) tile sums itself: total, ESS, its own exclusive
//       prefix; CDF of the tile, ancestor search (k3_tile), ll += max + log(mean w1)   :126-128
//   --  barrier (ancestors complete before the next gather)
//
// The per-particle arithmetic is the code of the three-launch path (propagate_particles,
// WeightSrc, k3_tile) and the sums are the same exact fixed-point integers, so both paths return
// the same bits (tests/test_gpu_parity.py::test_series_kernel_equals_three_launch_path).
//
// Memory visibility inside the launch: data written by another block in an earlier stage is read
// either with ld.global.cg (cloud, ancestors, tile tables) or with plain loads after the barrier's
// acquire fence, which invalidates L1 exactly as cooperative_groups::grid_group::sync() does; the
// non-coherent read-only path (ld.global.nc / __ldg) is never used for such data.
#pragma once
#include "cssm_kernels.cuh"

namespace cssm {

// per-observation record in device memory, in the filter dtype:
//   A[d] D[d] S[d] C[d]  y k0 k1 k2 k3 has_obs pad pad          (StepArgs without the padding)
constexpr int SERIES_REC_EXTRA = 8;
constexpr int SERIES_SMALL_MAX_TILES = 384;  // k_series_small: one tile per block, at most this many blocks (a multiple of 128)

// k_series_small: arrivals and the running max key side by side, read by ONE 16-byte load: a snapshot whose counter is
// complete carries the complete max (every block's RED.MAX precedes its releasing arrival), so the consumer needs no
// second round trip to fetch the key after the barrier.  Two pairs, by step parity.
struct __align__(16) SeriesPair {
  unsigned long long bar, key;
};
struct __align__(128) SeriesCtl {
  unsigned long long bar;       // k_series_multi: grid barrier arrivals (zero at launch)
  unsigned long long gkey[2];   // k_series_multi: ordered key of max(logw) by step parity (zero at launch)
  unsigned long long pad[13];
  SeriesPair pk[2];             // k_series_small (zero at launch)
  unsigned long long pad2[12];
};
static_assert(sizeof(SeriesCtl) == 256, "SeriesCtl layout");

struct SeriesArgs {
  void* x[2];               // ping-pong clouds; x[0] is the current one at entry
  void* logw;
  void* logw2;              // k_series_small: second log-weight buffer (steps alternate; the last step writes `logw`)
  int32_t* anc;
  unsigned long long* anc64;  // k_series_small: tagged ancestor words [Ns], zero at launch
  FilterScalars* sc;
  u128* tile_sum;           // [nt]
  u128* tile_q;             // [nt]
  double* tile_maxw;        // [nt]
  SeriesCtl* ctl;
  const void* recs;         // T records
  double* ll_steps;
  int* ess_steps;
  long long N, Ns;
  int T, d, nt, obs_kind;
  uint32_t key0, key1, step0;   // Philox step counter of the first step
  double inv_n;                 // 1/N when N is a power of two, else 0
  int tie_first;                // K3Ctl::tie_first
  unsigned long long* dbg;      // NULL, or 8 cycle counters per block [8][grid] (CSSM_SERIES_DEBUG): P1 B1 P2 B2 P3 anc-wait head
  Peers pr[2];                  // the (single-rank) topology with x[0] = the cloud read in even / odd steps: kept in
                                // the kernel's constant bank instead of a 400-byte struct in local memory
  unsigned long long* ll_rec;   // k_series_one: [T][6] exact totals, sum of squares and max key of every observed step
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release_gpu_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// All blocks of the (cooperative, fully resident) grid.  The arrival is a release at gpu scope --
// cumulative over the bar.sync before it, so it publishes the writes of the whole block -- and the
// poll is an acquire, after which the bar.sync hands the peers' writes to every thread of the block
// (ld.acquire.gpu also drops the SM's L1 lines, so later plain loads come from L2).  No stand-alone
// fence: membar.gl costs about a microsecond and a step has three barriers.
// Bounded: a block that waits longer than 2 s raises FLAG_COMM_TIMEOUT and every later barrier
// falls through, so a fault cannot hang the GPU.
__device__ __forceinline__ void grid_barrier(SeriesCtl* c, FilterScalars* sc, unsigned long long& target, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += G;
    red_release_gpu_add(&c->bar, 1ull);
    if (ld_acquire_gpu(&c->bar) < target && !(*(volatile int*)&sc->flags & FLAG_COMM_TIMEOUT)) {
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_gpu(&c->bar) < target) {
        if ((++spins & 1023u) == 0u) {
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 2000000000ull) {
            atomicOr(&sc->flags, FLAG_COMM_TIMEOUT);
            break;
          }
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void ld_acquire_gpu_pair(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// The same barrier on an (arrivals, key) pair: *key_out = the max key every block folded into the pair BEFORE arriving
// (thread 0 only; the caller broadcasts it).  Between the arrival and the first poll -- the barrier takes ~2.7 k cycles
// whatever the block does meanwhile -- every thread runs `idle_all` and thread 0 then `idle_t0`: work that does not
// depend on what the barrier publishes is free there.
// LEAD = false: the caller has already passed a block barrier behind the block's last global write of the stage (or
// thread 0 alone wrote), so thread 0 arrives at once.
template <bool LEAD = true, typename IdleAll, typename IdleT0>
__device__ __forceinline__ void grid_barrier_pair(SeriesPair* c, FilterScalars* sc, unsigned long long& target, unsigned G,
                                                  unsigned long long* key_out, IdleAll idle_all, IdleT0 idle_t0) {
  if (LEAD) __syncthreads();
  if (threadIdx.x == 0) {
    target += G;
    red_release_gpu_add(&c->bar, 1ull);
  }
  idle_all();
  if (threadIdx.x == 0) {
    idle_t0();
    unsigned long long n, k;
    ld_acquire_gpu_pair(&c->bar, n, k);
    if (n < target && !(*(volatile int*)&sc->flags & FLAG_COMM_TIMEOUT)) {
      unsigned long long t0 = 0;
      unsigned spins = 0;
      for (;;) {
        ld_acquire_gpu_pair(&c->bar, n, k);
        if (n >= target) break;
        if ((++spins & 1023u) == 0u) {
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 2000000000ull) {
            atomicOr(&sc->flags, FLAG_COMM_TIMEOUT);
            break;
          }
        }
      }
    }
    if (key_out != nullptr) *key_out = k;
  }
  __syncthreads();
}

// the kernels below are instantiated in cssm_series.cu only: the optimiser's choices for them and for the step kernels of
// cssm_api.cu then do not depend on each other; the host side asks for them by these two functions
void* series_small_kernel(int dtype, int items, int d, int resample_kind);
void* series_multi_kernel(int dtype, int items, int d, int resample_kind);

#define CSSM_STAMP(slot)                                              \
  if (sa.dbg != nullptr && threadIdx.x == 0) {                        \
    const long long now_ = clock64();                                 \
    s_dbg[slot] += (unsigned long long)(now_ - stamp_);               \
    stamp_ = now_;                                                    \
  }

// max over the warp of a thread's largest log-weight (a value of the filter dtype held in a double, never NaN): for fp32
// one REDUX.MAX on the order-preserving integer image instead of five shuffle rounds in fp64
template <typename real> __device__ __forceinline__ double warp_max_lw(double mx);
template <> __device__ __forceinline__ double warp_max_lw<double>(double mx) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  return mx;
}
template <> __device__ __forceinline__ double warp_max_lw<float>(double mx) {
  const unsigned b = __float_as_uint((float)mx);
  const unsigned k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  const unsigned r = __reduce_max_sync(0xffffffffu, k);
  return (double)__uint_as_float((r & 0x80000000u) ? (r & 0x7fffffffu) : ~r);
}

// the record of one observation -> the StepArgs in shared memory (one element per thread, rec_len <= TILE_THREADS)
template <typename real>
__device__ __forceinline__ void rec_to_args(StepArgs<real>& a, real rec_v, int d, int obs_kind) {
  const int i = threadIdx.x, e = i - 4 * d;
  if (i < 4 * d) {
    const int which = i / d, k = i - which * d;
    real* dst = (which == 0) ? a.A : (which == 1) ? a.D : (which == 2) ? a.S : a.C;
    dst[k] = rec_v;
  } else if (e == 0) a.y = rec_v;
  else if (e == 1) a.k0 = rec_v;
  else if (e == 2) a.k1 = rec_v;
  else if (e == 3) a.k2 = rec_v;
  else if (e == 4) a.k3 = rec_v;
  else if (e == 5) a.has_obs = (rec_v != (real)0) ? 1 : 0;
  else if (e == 6) { a.d = d; a.obs_kind = obs_kind; }
}

// k_series_small: TWO grid barriers per observed step.  The third one -- "every ancestor is written before the next
// gather" -- is replaced by tagged ancestor words: the search stores (step tag << 32 | index) and the thread that owns an
// output slot polls ITS OWN words until they carry the tag of the step.  What that removes from the hazards the barrier
// used to cover, and what covers them now:
//   * log-weights: the walk over a run of repeated keys (k3_tile) may read another tile's log-weights while that tile
//     already writes those of the next step -> two buffers, by step parity (the last step always uses sa.logw);
//   * the running max: (arrivals, key) pairs by step parity; block 0 clears the other pair between the two barriers of
//     a step, when nobody reads or updates it;
//   * ll / ESS of step s: block 0, inside the wait of the first barrier of step s + 1 (or after the loop);
//   * tile tables: written after the first barrier of step s + 1, which every block reaches after its search of step s.
// A step without an observation keeps one barrier (the clouds are ping-pong buffers).
// The normals of step s + 1 depend on nothing: every thread draws them while it waits for the second barrier of step s.
// ITEMS = 1 (256-particle tiles, two blocks per SM) while the cloud fits that way: the stages are chains of dependent
// instructions, and sixteen warps per SM with one particle per thread run them faster than eight with two.
template <typename real, int D, int KIND, int ITEMS>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_series_small(const __grid_constant__ SeriesArgs sa) {
  constexpr int TILE = TILE_THREADS * ITEMS;
  constexpr int NW = TILE_THREADS / 32;
  constexpr int NCH = (D > 0) ? (D + 3) / 4 : 1;  // chunks of four coordinates whose normals are drawn ahead
  typedef typename WeightSrc<real>::wt wt;
  __shared__ K3Smem<ITEMS> sm;
  __shared__ StepArgs<real> abuf[2];  // the constants of step s in abuf[s & 1]; the next step's are written a step ahead
  __shared__ double s_mx[NW], s_mxw[NW];
  __shared__ u128 s_r[3][NW], s_q2[NW];
  __shared__ u128 s_woff[NW + 1];  // exclusive prefix of the warp totals inside the tile, [NW] = the tile sum
  __shared__ int s_bad;
  __shared__ unsigned long long s_key;

  const int t = blockIdx.x;  // one tile per block, gridDim.x == nt
  const unsigned G = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d = (D > 0) ? D : sa.d;
  const long long N = sa.N, Ns = sa.Ns;
  unsigned long long target[2] = {0ull, 0ull};
  long long stamp_ = clock64();
  __shared__ unsigned long long s_dbg[16];  // CSSM_SERIES_DEBUG: cycles of thread 0 per stage (every block reports)
  if (threadIdx.x < 16) s_dbg[threadIdx.x] = 0ull;

  SumTables tb;
  tb.tile_sum = sa.tile_sum;
  tb.tile_maxw = sa.tile_maxw;
  tb.tile_q = sa.tile_q;
  tb.super_sum = nullptr;
  tb.super_q = nullptr;
  tb.super_ticket = nullptr;
  tb.nt = sa.nt;
  tb.ns = 0;

  int cur = 0;
  unsigned anc_tag = 0;  // != 0: the ancestors of the previous step, as tagged words
  const int rec_len = 4 * d + SERIES_REC_EXTRA;  // <= 136 <= TILE_THREADS: one element per thread
  {
    const real r0 = ((int)threadIdx.x < rec_len) ? __ldg(reinterpret_cast<const real*>(sa.recs) + threadIdx.x) : (real)0;
    rec_to_args<real>(abuf[0], r0, d, sa.obs_kind);
    if (threadIdx.x == 0) s_bad = 0;
  }
  // the record of step s + 1 travels in a register during step s and is stored into the other buffer at the top of it
  real rec_v = (sa.T > 1 && (int)threadIdx.x < rec_len) ? __ldg(reinterpret_cast<const real*>(sa.recs) + rec_len + threadIdx.x) : (real)0;
  const long long i0 = (long long)t * TILE + (long long)threadIdx.x * ITEMS;
  real zpre[NCH][4][ITEMS];
  bool z_ready = false;
  K3Ctl kc;
  kc.parity = 0;
  kc.obs_seq = 0;
  kc.gstep = 0;
  kc.inv_n = sa.inv_n;
  kc.direct = 0;
  kc.add_ll = 1;
  kc.use_u_inj = 0;
  kc.tie_first = sa.tie_first;
  kc.defer_ll = 1;
  kc.fast_ok = 0;
  kc.key0 = sa.key0;
  kc.key1 = sa.key1;
  kc.ll_steps = sa.ll_steps;
  kc.ess_steps = sa.ess_steps;
  kc.anc64 = sa.anc64;
  kc.dbg = sa.dbg != nullptr ? s_dbg : nullptr;
  __syncthreads();
  // ---- the accountant: one extra block (the last of the grid) that owns no particles.  It takes part in every barrier
  //      -- never as the last to arrive -- and does what one thread has to do once per step: ll += max + log(mean w1) and
  //      the ESS from the exact sums (three divisions and a logarithm in fp64), the reset of the other parity's max.  In
  //      a block that also carries a tile that serial work made the block late for the next barrier, and with it the grid.
  if (t == sa.nt) {
    for (int s = 0; s < sa.T; ++s) {
      const int pp = s & 1;
      SeriesPair* const pk = &sa.ctl->pk[pp];
      const bool obs = __ldg(reinterpret_cast<const real*>(sa.recs) + (size_t)s * rec_len + 4 * d + 5) != (real)0;
      if (!obs) {
        if (threadIdx.x == 0) sa.ctl->pk[pp ^ 1].key = 0ull;
        grid_barrier_pair(pk, sa.sc, target[pp], G, nullptr, [] {}, [] {});
        continue;
      }
      grid_barrier_pair(pk, sa.sc, target[pp], G, &s_key, [] {}, [] {});
      const unsigned long long key = s_key;
      if (threadIdx.x == 0) sa.ctl->pk[pp ^ 1].key = 0ull;  // nobody reads or updates the other pair between the two barriers
      grid_barrier_pair(pk, sa.sc, target[pp], G, nullptr, [] {}, [] {});
      u128 a1 = make_u128(0, 0), a2 = make_u128(0, 0);
      for (int tt = threadIdx.x; tt < sa.nt; tt += TILE_THREADS) {
        const ulonglong2 v1 = __ldcg(reinterpret_cast<const ulonglong2*>(&sa.tile_sum[tt]));
        const ulonglong2 v2 = __ldcg(reinterpret_cast<const ulonglong2*>(&sa.tile_q[tt]));
        a1 = add128(a1, make_u128(v1.x, v1.y));
        a2 = add128(a2, make_u128(v2.x, v2.y));
      }
      a1 = warp_sum128(a1);
      a2 = warp_sum128(a2);
      if (lane == 0) { s_r[0][wid] = a1; s_r[1][wid] = a2; }
      __syncthreads();
      if (threadIdx.x == 0) {
        u128 tot = s_r[0][0], qsum = s_r[1][0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          tot = add128(tot, s_r[0][w]);
          qsum = add128(qsum, s_r[1][w]);
        }
        kc.step = sa.step0 + (uint32_t)s;
        kc.step_slot = s;
        ll_ess_update<real, false>(sa.sc, kc, tot, qsum, key, (long long)N, false);
      }
      __syncthreads();  // s_r, s_key
    }
    return;
  }
  for (int s = 0; s < sa.T; ++s) {
    const StepArgs<real>& a = abuf[s & 1];
    // the constants of step s + 1 go into the other buffer -- last read in step s - 1 -- inside the wait of this step's
    // first barrier, and the record of step s + 2 is requested there
    auto next_consts = [&]() {
      if (s + 1 < sa.T) rec_to_args<real>(abuf[(s + 1) & 1], rec_v, d, sa.obs_kind);
      if (s + 2 < sa.T && (int)threadIdx.x < rec_len) rec_v = __ldg(reinterpret_cast<const real*>(sa.recs) + (size_t)(s + 2) * rec_len + threadIdx.x);
    };
    const int has_obs = a.has_obs;
    const uint32_t step = sa.step0 + (uint32_t)s;
    const int pp = s & 1;
    SeriesPair* const pk = &sa.ctl->pk[pp];
    real* const logw = reinterpret_cast<real*>(((sa.T - 1 - s) & 1) ? sa.logw2 : sa.logw);

    CSSM_STAMP(6)
    // ---- P1 ------------------------------------------------------------------------------------
    const Peers& pr = sa.pr[cur];
    double mx;
    bool bad;
    real lw[ITEMS];
    int sidx[ITEMS];
    if (anc_tag != 0u) {  // the ancestors of this thread's two slots: poll until both words carry the tag of the last search
      unsigned long long w[ITEMS];
      auto load = [&]() {
        if (ITEMS == 2) ld_gpu_pair(&sa.anc64[i0], w[0], w[ITEMS - 1]);
        else w[0] = ld_gpu(&sa.anc64[i0]);
      };
      auto complete = [&]() {  // slots past the end of a ragged cloud are never written
        bool ok = true;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) ok &= !(i0 + j < N) || (unsigned)(w[j] >> 32) == anc_tag;
        return ok;
      };
      load();
      if (!complete()) {
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
          load();
          if (complete()) break;
          if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull || (*(volatile int*)&sa.sc->flags & FLAG_COMM_TIMEOUT)) {
              atomicOr(&sa.sc->flags, FLAG_COMM_TIMEOUT);
#pragma unroll
              for (int j = 0; j < ITEMS; ++j) w[j] = (unsigned long long)(i0 + j);  // indices inside the cloud; the run is reported as failed
              break;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) sidx[j] = (int)(unsigned)w[j];
    }
    CSSM_STAMP(5)
    propagate_particles<real, D, ITEMS, true>(a, pr, reinterpret_cast<real*>(sa.x[cur ^ 1]), nullptr, logw, nullptr, N, Ns, 0ull,
                                              sa.key0, sa.key1, step, i0, mx, bad, lw, anc_tag != 0u ? sidx : nullptr,
                                              (D > 0 && z_ready) ? zpre : nullptr);
    cur ^= 1;
    anc_tag = 0u;
    z_ready = false;
    if (!has_obs) {  // propagated only (:121); one barrier: the clouds are ping-pong buffers
      grid_barrier_pair(pk, sa.sc, target[pp], G, nullptr, next_consts, [] {});
      continue;
    }
    mx = warp_max_lw<real>(mx);
    if (lane == 0) s_mx[wid] = mx;
    if (bad) s_bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m2 = s_mx[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) m2 = fmax(m2, s_mx[w]);
      atomicMax(&pk->key, ord_key(m2));
      if (s_bad) { atomicOr(&sa.sc->flags, FLAG_NAN_WEIGHT); s_bad = 0; }
    }
    CSSM_STAMP(0)
    kc.step = step;
    kc.step_slot = s;
    kc.anc_tag = (unsigned)s + 1u;
    // the max arrives with the barrier's own snapshot; in its shadow: the next step's constants, the resampling uniform
    grid_barrier_pair(pk, sa.sc, target[pp], G, &s_key,
                             [&] { next_consts(); if (threadIdx.x == 64) k3_prepare<ITEMS>(sm, sa.sc, kc); },
                             [] {});
    CSSM_STAMP(1)

    // ---- P2: weights, and the LOCAL part of the tile scan (its by-product is the exact tile sum) ------------
    const unsigned long long key = s_key;
    const PreScan ps = pre_scan(key, false);
    const int qb = ps.qb;
    WeightSrc<real> ws{logw, nullptr, ps.gmax};
    TileScan<real, ITEMS> scan;  // this thread's weights, from the log-weights it still holds in registers
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) scan.w[j] = (i0 + j < N) ? ws.weight(lw[j]) : (wt)0;
    {
      if (threadIdx.x == 32) sm.s_wnext = (t < sa.nt - 1) ? (double)ws((long long)(t + 1) * TILE) : 0.0;  // first weight of the next tile
      u128 acc2 = make_u128(0, 0);
      wt mxv = (wt)0;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        WeightSrc<real>::acc_sq(acc2, scan.w[j], 1.0);
        mxv = scan.w[j] > mxv ? scan.w[j] : mxv;
      }
      const double mxw = WeightSrc<real>::warp_max(mxv);
      acc2 = WeightSrc<real>::warp_sum_sq(acc2);
      tile_scan_local<real, ITEMS>(qb, i0, N, sm.s_warp, sm.s_minw, scan);
      if (lane == 0) { s_q2[wid] = acc2; s_mxw[wid] = mxw; }
      __syncthreads();
      if (threadIdx.x == 0) {
        u128 t1 = sm.s_warp[0], t2 = s_q2[0];
        double m2 = s_mxw[0];
        s_woff[0] = make_u128(0, 0);
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          s_woff[w] = t1;  // exact sum of the warps before warp w: the scan's finish reads one word instead of adding eight
          t1 = add128(t1, sm.s_warp[w]);
          t2 = add128(t2, s_q2[w]);
          m2 = fmax(m2, s_mxw[w]);
        }
        s_woff[NW] = t1;
        sa.tile_sum[t] = t1;
        sa.tile_q[t] = t2;
        sa.tile_maxw[t] = m2;
      }
    }
    CSSM_STAMP(2)
    // in the shadow of the second barrier: the normals of the next step (they depend on nothing but the slot and the step)
    grid_barrier_pair(pk, sa.sc, target[pp], G, nullptr,
                      [&] {
                        if (D > 0 && s + 1 < sa.T) {
#pragma unroll
                          for (int c = 0; c < NCH; ++c) chunk_noise<real, ITEMS>(d, 4 * c, 0ull, i0, step + 1u, sa.key0, sa.key1, zpre[c]);
                          z_ready = true;
                        }
                      },
                      [] {});
    CSSM_STAMP(3)

    // ---- P3: totals and this tile's exclusive prefix from the tile sums, then the FINISH of the scan + the search ----
    {
      if (sa.dbg != nullptr && threadIdx.x == 0) s_dbg[7] = (unsigned long long)clock64();
      // warps 0-3: the total, 32 tiles of every 128 each; warps 4-7 likewise the tiles before this one.  One 16-byte L2
      // load per lane and chunk, issued back to back; one warp-wide sum per warp; eight words through shared memory.
      {
        const int lim = (wid >= 4) ? t : sa.nt;
        u128 a_ = make_u128(0, 0);
#pragma unroll
        for (int c = 0; c < SERIES_SMALL_MAX_TILES / 128; ++c) {
          const int tt = c * 128 + (wid & 3) * 32 + lane;
          if (tt < lim) {
            const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(&sa.tile_sum[tt]));
            a_ = add128(a_, make_u128(v.x, v.y));
          }
        }
        a_ = warp_sum128(a_);
        if (lane == 0) s_r[0][wid] = a_;
      }
      __syncthreads();
      const u128 tot = add128(add128(s_r[0][0], s_r[0][1]), add128(s_r[0][2], s_r[0][3]));
      const u128 excl = add128(add128(s_r[0][4], s_r[0][5]), add128(s_r[0][6], s_r[0][7]));
      const u128 qsum = make_u128(0, 0);  // the sum of squares only feeds the ESS: the accountant's
      k3_tile<real, ITEMS, KIND, false>(sm, logw, nullptr, N, sa.sc, tb, pr, kc, nullptr, nullptr, t, tot, qsum, key, excl, &scan, s_woff);
      anc_tag = kc.anc_tag;
      // no block barrier here: what the next step writes before its first barrier (abuf of step s + 2, s_mx) is not read
      // by the search, and the search's own shared memory is next written behind that barrier
      CSSM_STAMP(4)
    }
  }
  if (sa.dbg != nullptr && threadIdx.x == 0)
    for (int k = 0; k < 16; ++k) sa.dbg[(size_t)k * sa.nt + t] = s_dbg[k];
  // the host's view: plain int32 ancestors of the last search
  if (anc_tag != 0u) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (i0 + j >= N) continue;
      unsigned long long w = ld_gpu(&sa.anc64[i0 + j]);
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while ((unsigned)(w >> 32) != anc_tag) {
        w = ld_gpu(&sa.anc64[i0 + j]);
        if ((++spins & 1023u) == 0u) {
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 2000000000ull) { atomicOr(&sa.sc->flags, FLAG_COMM_TIMEOUT); w = (unsigned long long)(i0 + j); break; }
        }
      }
      sa.anc[i0 + j] = (int32_t)(unsigned)w;
    }
  }
}
#undef CSSM_STAMP

// ---------------------------------------------------------------------------------------------
// TINY clouds -- the particle counts of the reference's own examples (100 .. 1000 particles,
// examples/DetermineParameters.scala:70, examples/Filtering.scala:24): the whole llFilter in ONE BLOCK.  No grid-wide
// exchange exists: the max is a block reduction, the tile sum IS the total, the exclusive prefix is zero, ancestors go
// through L2 between two block barriers.  Same per-particle code (propagate_particles, WeightSrc, tile_scan_*, k3_tile),
// same exact sums: bit-identical to the other schedules.  256 threads x 4 particles.
// ---------------------------------------------------------------------------------------------
template <> struct VecN<double, 4> { typedef double4 type; typedef int4 itype; };
constexpr int SERIES_ONE_ITEMS = 4;
constexpr int SERIES_ONE_MAX_N = TILE_THREADS * SERIES_ONE_ITEMS;
void* series_one_kernel(int dtype, int d, int resample_kind);

template <typename real, int D, int KIND>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_series_one(const __grid_constant__ SeriesArgs sa) {
  constexpr int ITEMS = SERIES_ONE_ITEMS;
  constexpr int NW = TILE_THREADS / 32;
  typedef typename WeightSrc<real>::wt wt;
  __shared__ K3Smem<ITEMS> sm;
  __shared__ StepArgs<real> abuf[2];  // the constants of step s in abuf[s & 1]; the next step's are written a step ahead
  __shared__ double s_mx[NW], s_mxw[NW];
  __shared__ u128 s_q2[NW], s_woff[NW + 1], s_qsum;
  __shared__ int s_bad;

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d = (D > 0) ? D : sa.d;
  const long long N = sa.N, Ns = sa.Ns;
  real* const logw = reinterpret_cast<real*>(sa.logw);
  SumTables tb;
  tb.tile_sum = sa.tile_sum;
  tb.tile_maxw = sa.tile_maxw;
  tb.tile_q = sa.tile_q;
  tb.super_sum = nullptr;
  tb.super_q = nullptr;
  tb.super_ticket = nullptr;
  tb.nt = 1;
  tb.ns = 0;
  K3Ctl kc;
  kc.parity = 0;
  kc.obs_seq = 0;
  kc.gstep = 0;
  kc.inv_n = sa.inv_n;
  kc.direct = 0;
  kc.add_ll = 1;
  kc.use_u_inj = 0;
  kc.tie_first = sa.tie_first;
  kc.defer_ll = 1;  // ll / ESS of all steps at the end of the kernel, in parallel over the steps (below)
  kc.fast_ok = 0;
  kc.key0 = sa.key0;
  kc.key1 = sa.key1;
  kc.ll_steps = sa.ll_steps;
  kc.ess_steps = sa.ess_steps;
  kc.anc64 = nullptr;
  kc.anc_tag = 0;
  kc.dbg = nullptr;

  int cur = 0;
  bool anc_valid = false;
  const int rec_len = 4 * d + SERIES_REC_EXTRA;
  const long long i0 = (long long)threadIdx.x * ITEMS;
  if (threadIdx.x == 0) s_bad = 0;
  {
    const real r0 = ((int)threadIdx.x < rec_len) ? __ldg(reinterpret_cast<const real*>(sa.recs) + threadIdx.x) : (real)0;
    rec_to_args<real>(abuf[0], r0, d, sa.obs_kind);
  }
  // the record of step s + 1 travels in a register during step s and is stored into the other buffer at the top of it
  real rec_v = (sa.T > 1 && (int)threadIdx.x < rec_len) ? __ldg(reinterpret_cast<const real*>(sa.recs) + rec_len + threadIdx.x) : (real)0;
  for (int s = 0; s < sa.T; ++s) {
    __syncthreads();  // the step before is complete in every thread: its ancestors and cloud, and abuf[(s + 1) & 1] is free
    const StepArgs<real>& a = abuf[s & 1];
    if (s + 1 < sa.T) rec_to_args<real>(abuf[(s + 1) & 1], rec_v, d, sa.obs_kind);
    if (s + 2 < sa.T && (int)threadIdx.x < rec_len) rec_v = __ldg(reinterpret_cast<const real*>(sa.recs) + (size_t)(s + 2) * rec_len + threadIdx.x);
    const int has_obs = a.has_obs;
    const uint32_t step = sa.step0 + (uint32_t)s;
    // ---- P1 ----
    const Peers& pr = sa.pr[cur];
    double mx;
    bool bad;
    real lw[ITEMS];
    propagate_particles<real, D, ITEMS, true>(a, pr, reinterpret_cast<real*>(sa.x[cur ^ 1]), anc_valid ? sa.anc : nullptr, logw, nullptr,
                                              N, Ns, 0ull, sa.key0, sa.key1, step, i0, mx, bad, lw);
    cur ^= 1;
    anc_valid = false;
    if (!has_obs) continue;  // the barrier at the top of the next step orders the clouds
    mx = warp_max_lw<real>(mx);
    if (lane == 0) s_mx[wid] = mx;
    if (bad) s_bad = 1;
    __syncthreads();
    double m2 = s_mx[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m2 = fmax(m2, s_mx[w]);
    const unsigned long long key = ord_key(m2);
    if (threadIdx.x == 0 && s_bad) { atomicOr(&sa.sc->flags, FLAG_NAN_WEIGHT); s_bad = 0; }
    // ---- P2 ----
    const PreScan ps = pre_scan(key, false);
    const int qb = ps.qb;
    WeightSrc<real> ws{logw, nullptr, ps.gmax};
    kc.step = step;
    kc.step_slot = s;
    TileScan<real, ITEMS> scan;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) scan.w[j] = (i0 + j < N) ? ws.weight(lw[j]) : (wt)0;
    {
      if (threadIdx.x == 64) k3_prepare<ITEMS>(sm, sa.sc, kc);
      if (threadIdx.x == 32) sm.s_wnext = 0.0;  // no tile behind this one
      u128 acc2 = make_u128(0, 0);
      wt mxv = (wt)0;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        WeightSrc<real>::acc_sq(acc2, scan.w[j], 1.0);
        mxv = scan.w[j] > mxv ? scan.w[j] : mxv;
      }
      const double mxw = WeightSrc<real>::warp_max(mxv);
      acc2 = WeightSrc<real>::warp_sum_sq(acc2);
      tile_scan_local<real, ITEMS>(qb, i0, N, sm.s_warp, sm.s_minw, scan);
      if (lane == 0) { s_q2[wid] = acc2; s_mxw[wid] = mxw; }
      __syncthreads();
      if (threadIdx.x == 0) {
        u128 t1 = sm.s_warp[0], t2 = s_q2[0];
        double mw = s_mxw[0];
        s_woff[0] = make_u128(0, 0);
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          s_woff[w] = t1;
          t1 = add128(t1, sm.s_warp[w]);
          t2 = add128(t2, s_q2[w]);
          mw = fmax(mw, s_mxw[w]);
        }
        s_woff[NW] = t1;
        s_qsum = t2;
        sa.tile_sum[0] = t1;  // the tables a walk over a run of repeated keys would read (it cannot leave the only tile)
        sa.tile_q[0] = t2;
        sa.tile_maxw[0] = mw;
        unsigned long long* r = sa.ll_rec + (size_t)s * 6;
        r[0] = t1.lo; r[1] = t1.hi; r[2] = t2.lo; r[3] = t2.hi; r[4] = key; r[5] = 1ull;
      }
      __syncthreads();
    }
    // ---- P3: the tile sum is the total, nothing lies before the tile ----
    k3_tile<real, ITEMS, KIND, false>(sm, logw, nullptr, N, sa.sc, tb, pr, kc, nullptr, nullptr, 0, s_woff[NW], s_qsum, key,
                                      make_u128(0, 0), &scan, s_woff);
    anc_valid = true;
  }
  // ---- ll and ESS of every observed step.  One thread per step evaluates max + log(mean w1) and floor(1 / sum wn^2)
  //      (three divisions and a logarithm in fp64: inside the loop they held one warp back by ~2 k cycles per step and
  //      the block with it); thread 0 then adds the increments in order -- the same sums, the same bits. ----
  __syncthreads();
  for (int s = threadIdx.x; s < sa.T; s += TILE_THREADS) {
    unsigned long long* r = sa.ll_rec + (size_t)s * 6;
    if (r[5] == 0ull) continue;  // no observation
    double incr;
    int ess, flags;
    ll_ess_terms<real>(make_u128(r[0], r[1]), make_u128(r[2], r[3]), r[4], (long long)N, incr, ess, flags);
    r[2] = (unsigned long long)__double_as_longlong(incr);  // the sum of squares has been used
    r[3] = ((unsigned long long)(unsigned)flags << 32) | (unsigned long long)(unsigned)ess;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ll = sa.sc->ll;
    int ess = sa.sc->ess, flags = 0, last = -1;
    for (int s = 0; s < sa.T; ++s) {
      const unsigned long long* r = sa.ll_rec + (size_t)s * 6;
      if (r[5] == 0ull) continue;
      ll = ll + __longlong_as_double((long long)r[2]);
      ess = (int)(unsigned)(r[3] & 0xffffffffull);
      flags |= (int)(unsigned)(r[3] >> 32);
      if (sa.ll_steps) sa.ll_steps[s] = ll;
      if (sa.ess_steps) sa.ess_steps[s] = ess;
      last = s;
    }
    if (last >= 0) {
      const unsigned long long* r = sa.ll_rec + (size_t)last * 6;
      const PreScan ps = pre_scan(r[4], false);
      sa.sc->gmax = ps.gmax;
      sa.sc->total = dbl128(make_u128(r[0], r[1]), ps.qb);
      sa.sc->qb = ps.qb;
      sa.sc->ll_incr = __longlong_as_double((long long)r[2]);
      sa.sc->ll = ll;
      sa.sc->ess = ess;
      if (flags) atomicOr(&sa.sc->flags, flags);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same single-launch schedule for MID-SIZE clouds (up to a few million particles): more tiles
// than resident blocks, so block b owns the tiles b, b+G, b+2G, ... and loops over them in every
// stage; log-weights go through L2 between the stages instead of staying in registers.  At 2^20
// particles a stage is 5-15 us of work and the launch / drain gaps of the three-launch step cost a
// third of the step; here they shrink to three grid barriers.  Same per-particle code, same exact
// sums: bit-identical to the other two schedules.
// ---------------------------------------------------------------------------------------------
template <typename real, int D, int KIND, int ITEMS>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_series_multi(const __grid_constant__ SeriesArgs sa) {
  constexpr int TILE = TILE_THREADS * ITEMS;
  constexpr int PPT = VecOf<real>::PPT;
  constexpr int CHUNK = TILE_THREADS * PPT;
  typedef typename WeightSrc<real>::wt wt;
  // the step constants are only read in P1 and the scan + search memory only in P3: they share storage
  __shared__ union SmemU { K3Smem<ITEMS> sm; StepArgs<real> a; } smem_u;
  K3Smem<ITEMS>& sm = smem_u.sm;
  StepArgs<real>& a = smem_u.a;
  __shared__ double s_mx[TILE_THREADS / 32], s_mxw[TILE_THREADS / 32];
  __shared__ u128 s_r[3][TILE_THREADS / 32];
  __shared__ u128 s_excl_run;
  __shared__ int s_bad;

  const int b = blockIdx.x;
  const unsigned G = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d = (D > 0) ? D : sa.d;
  const long long N = sa.N, Ns = sa.Ns;
  const long long n_chunks = (N + CHUNK - 1) / CHUNK;
  real* const logw = reinterpret_cast<real*>(sa.logw);
  unsigned long long target = 0;

  SumTables tb;
  tb.tile_sum = sa.tile_sum;
  tb.tile_maxw = sa.tile_maxw;
  tb.tile_q = sa.tile_q;
  tb.super_sum = nullptr;
  tb.super_q = nullptr;
  tb.super_ticket = nullptr;
  tb.nt = sa.nt;
  tb.ns = 0;

  int cur = 0, n_obs = 0;
  bool anc_valid = false;
  const int rec_len = 4 * d + SERIES_REC_EXTRA;
  real rec_v = ((int)threadIdx.x < rec_len) ? __ldg(reinterpret_cast<const real*>(sa.recs) + threadIdx.x) : (real)0;
  for (int s = 0; s < sa.T; ++s) {
    {
      const int i = threadIdx.x, e = i - 4 * d;
      if (i < 4 * d) {
        const int which = i / d, k = i - which * d;
        real* dst = (which == 0) ? a.A : (which == 1) ? a.D : (which == 2) ? a.S : a.C;
        dst[k] = rec_v;
      } else if (e == 0) a.y = rec_v;
      else if (e == 1) a.k0 = rec_v;
      else if (e == 2) a.k1 = rec_v;
      else if (e == 3) a.k2 = rec_v;
      else if (e == 4) a.k3 = rec_v;
      else if (e == 5) a.has_obs = (rec_v != (real)0) ? 1 : 0;
      else if (e == 6) { a.d = d; a.obs_kind = sa.obs_kind; s_bad = 0; }
      if (s + 1 < sa.T && i < rec_len) rec_v = __ldg(reinterpret_cast<const real*>(sa.recs) + (size_t)(s + 1) * rec_len + i);
    }
    __syncthreads();
    const int has_obs = a.has_obs;
    const int par = n_obs & 1;
    const uint32_t step = sa.step0 + (uint32_t)s;

    // ---- P1: chunks b, b+G, ... of 256*PPT particles --------------------------------------------------
    const Peers& pr = sa.pr[cur];
    double mx = -__longlong_as_double(0x7FF0000000000000ll);
    bool bad = false;
    for (long long c = b; c < n_chunks; c += G) {
      const long long i0 = c * CHUNK + (long long)threadIdx.x * PPT;
      double mxc;
      bool badc;
      if ((c + 1) * CHUNK <= N)
        propagate_particles<real, D, PPT, true, true>(a, pr, reinterpret_cast<real*>(sa.x[cur ^ 1]), anc_valid ? sa.anc : nullptr,
                                                      logw, nullptr, N, Ns, 0ull, sa.key0, sa.key1, step, i0, mxc, badc);
      else
        propagate_particles<real, D, PPT, true, false>(a, pr, reinterpret_cast<real*>(sa.x[cur ^ 1]), anc_valid ? sa.anc : nullptr,
                                                       logw, nullptr, N, Ns, 0ull, sa.key0, sa.key1, step, i0, mxc, badc);
      mx = fmax(mx, mxc);
      bad |= badc;
    }
    cur ^= 1;
    anc_valid = false;
    if (!has_obs) {
      grid_barrier(sa.ctl, sa.sc, target, G);
      continue;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if (lane == 0) s_mx[wid] = mx;
    if (bad) s_bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m2 = s_mx[0];
      for (int w = 1; w < TILE_THREADS / 32; ++w) m2 = fmax(m2, s_mx[w]);
      atomicMax(&sa.ctl->gkey[par], ord_key(m2));
      if (s_bad) atomicOr(&sa.sc->flags, FLAG_NAN_WEIGHT);
    }
    grid_barrier(sa.ctl, sa.sc, target, G);

    // ---- P2: exact sums of the tiles b, b+G, ... ------------------------------------------------------
    const unsigned long long key = ld_gpu(&sa.ctl->gkey[par]);
    const PreScan ps = pre_scan(key, false);
    const int qb = ps.qb;
    WeightSrc<real> ws{logw, nullptr, ps.gmax};
    if (b == 0 && threadIdx.x == 0) sa.ctl->gkey[par ^ 1] = 0ull;
    for (int t = b; t < sa.nt; t += G) {
      wt wv[ITEMS];
      ws.template load<ITEMS>((long long)t * TILE + threadIdx.x, TILE_THREADS, N, wv);
      u128 acc = make_u128(0, 0), acc2 = make_u128(0, 0);
      wt mxv = (wt)0;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        acc = add128(acc, WeightSrc<real>::fix(wv[j], qb));
        WeightSrc<real>::acc_sq(acc2, wv[j], 1.0);
        mxv = wv[j] > mxv ? wv[j] : mxv;
      }
      const double mxw = WeightSrc<real>::warp_max(mxv);
      acc = warp_sum128(acc);
      acc2 = WeightSrc<real>::warp_sum_sq(acc2);
      __syncthreads();  // s_r / s_mxw of the previous tile have been read
      if (lane == 0) { s_r[0][wid] = acc; s_r[1][wid] = acc2; s_mxw[wid] = mxw; }
      __syncthreads();
      if (threadIdx.x == 0) {
        u128 t1 = s_r[0][0], t2 = s_r[1][0];
        double m2 = s_mxw[0];
        for (int w = 1; w < TILE_THREADS / 32; ++w) {
          t1 = add128(t1, s_r[0][w]);
          t2 = add128(t2, s_r[1][w]);
          m2 = fmax(m2, s_mxw[w]);
        }
        sa.tile_sum[t] = t1;
        sa.tile_q[t] = t2;
        sa.tile_maxw[t] = m2;
      }
    }
    grid_barrier(sa.ctl, sa.sc, target, G);

    // ---- P3: totals once per block, then scan + search of the tiles b, b+G, ... -----------------------
    {
      u128 at = make_u128(0, 0), aq = make_u128(0, 0), ae = make_u128(0, 0);
      for (int tt = threadIdx.x; tt < sa.nt; tt += TILE_THREADS) {
        const u128 v = ld_gpu128(&sa.tile_sum[tt]);
        at = add128(at, v);
        if (tt < b) ae = add128(ae, v);
        aq = add128(aq, ld_gpu128(&sa.tile_q[tt]));
      }
      at = warp_sum128(at);
      aq = warp_sum128(aq);
      ae = warp_sum128(ae);
      if (lane == 0) { s_r[0][wid] = at; s_r[1][wid] = aq; s_r[2][wid] = ae; }
      __syncthreads();
      if (threadIdx.x == 0) {
        u128 r0 = s_r[0][0], r1 = s_r[1][0], r2 = s_r[2][0];
        for (int w = 1; w < TILE_THREADS / 32; ++w) {
          r0 = add128(r0, s_r[0][w]);
          r1 = add128(r1, s_r[1][w]);
          r2 = add128(r2, s_r[2][w]);
        }
        sm.s_tot = r0;
        sm.s_q = r1;
        s_excl_run = r2;
      }
      __syncthreads();
      const u128 tot = sm.s_tot, qsum = sm.s_q;
      K3Ctl kc;
      kc.anc64 = nullptr;
      kc.anc_tag = 0;
      kc.dbg = nullptr;
      kc.fast_ok = 0;
      kc.parity = 0;
      kc.obs_seq = 0;
      kc.gstep = 0;
      kc.inv_n = sa.inv_n;
      kc.direct = 0;
      kc.add_ll = 1;
      kc.use_u_inj = 0;
      kc.tie_first = sa.tie_first;
      kc.defer_ll = 0;
      kc.key0 = sa.key0;
      kc.key1 = sa.key1;
      kc.step = step;
      kc.ll_steps = sa.ll_steps;
      kc.ess_steps = sa.ess_steps;
      kc.step_slot = s;
      for (int t = b; t < sa.nt; t += G) {
        const u128 excl = s_excl_run;
        __syncthreads();  // everyone holds excl (and is done with the shared memory of the previous tile)
        k3_tile<real, ITEMS, KIND, false>(sm, logw, nullptr, N, sa.sc, tb, pr, kc, nullptr, nullptr, t, tot, qsum, key, excl);
        if (t + (int)G < sa.nt) {  // exclusive prefix of the block's next tile: add the tiles t .. t+G-1
          u128 ad = make_u128(0, 0);
          for (int tt = t + threadIdx.x; tt < t + (int)G; tt += TILE_THREADS) ad = add128(ad, ld_gpu128(&sa.tile_sum[tt]));
          ad = warp_sum128(ad);
          __syncthreads();
          if (lane == 0) s_r[0][wid] = ad;
          __syncthreads();
          if (threadIdx.x == 0) {
            u128 r0 = s_excl_run;
            for (int w = 0; w < TILE_THREADS / 32; ++w) r0 = add128(r0, s_r[0][w]);
            s_excl_run = r0;
          }
          __syncthreads();
        }
      }
      anc_valid = true;
      ++n_obs;
    }
    grid_barrier(sa.ctl, sa.sc, target, G);
  }
}

}  // namespace cssm
