// cssm_kernels.cuh -- the sm_100a kernels of the particle-filter hot path.
//
//   k_init_particles      K0  Sde.initialState                model/Sde.scala:75-80,104-108,152-156
//   k_propagate_weight    K1  stepFunction + f + dataLikelihood + running max, with the gather of
//                             the previous resampling fused into the load
//                                                              model/ParticleFilter.scala:118,123-124
//   k_lgcp_weight         K1' FilterLgcp.calcWeight            model/ParticleFilter.scala:184-217
//   k_weight_total        K2a sum of w1 = exp(logw - max)      model/ParticleFilter.scala:125, :522-524
//   k_tile_sums           K2b normalise + per-tile sums + ESS  model/Resampling.scala:21-24, PF :431-434
//   k_scan_tiles          K2c tile prefix, ll increment, ESS   model/ParticleFilter.scala:127-128
//   k_scan_search         K3+K4 inclusive CDF + ancestor search (systematic / stratified)
//                                                              model/Resampling.scala:36-86
//   k_multinomial_search  K4' per-draw inverse CDF             model/Resampling.scala:92-96
//   k_gather              K5  particles(ancestor)              model/Resampling.scala:42,95
//
// Layout: structure of arrays, x[k*Ns + i] (Ns = N rounded up to 64) -- every access below is
// unit-stride across a warp except the ancestor gather, whose indices are non-decreasing for
// systematic/stratified resampling.  All weight sums are exact 2^-96 fixed point (cssm_common.cuh),
// so no result depends on the launch geometry.
#pragma once
#include "cssm_common.cuh"
#include "../../include/cssm.h"

namespace cssm {

constexpr int MAXD = 32;          // by-value kernel argument budget (5*32*8 B = 1.25 KB)
constexpr int TILE = 2048;        // elements per scan tile
constexpr int TILE_THREADS = 256; // 8 elements per thread
constexpr int TILE_ITEMS = TILE / TILE_THREADS;

// per-observation constants, built on the host in fp64 and rounded to the filter dtype.
// transition of coordinate k:  x' = A*(x - M) + M + D + S*z     (exact or Euler-Maruyama)
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

// device-resident scalars of one filter
struct Scalars {
  unsigned long long gmax_key;  // ordered key of max(logw), atomicMax target (0 = below everything)
  u128 tot;                     // sum fix(w1, qb)
  u128 ess_acc;                 // sum fix(wn^2, 96)
  double gmax, total, u, ll, ll_incr;
  int ess, flags, qb, pad;
  double u_inj;                 // injected systematic uniform
  unsigned long long n_launch_dummy;
};
enum : int { FLAG_NAN_WEIGHT = 1, FLAG_ZERO_TOTAL = 2, FLAG_CLAMPED = 4 };

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
template <typename real>
__device__ __forceinline__ real obs_loglik(const StepArgs<real>& a, real g) {
  switch (a.obs_kind) {
    case CSSM_OBS_POISSON:
      return r_fma<real>(a.k0, g, -r_exp<real>(g)) - a.k1;  // -mean + k*log(mean) - lgamma(k+1), log(exp(g)) = g
    case CSSM_OBS_NEGBIN: {
      real L = r_log<real>(r_exp<real>(g) + a.k1);  // log(mu + size)
      return a.k2 + a.k1 * (a.k3 - L) + a.k0 * (g - L);
    }
    case CSSM_OBS_NORMAL: {
      real dd = (a.y - g) / a.k0;
      return -dd * dd / (real)2 - a.k1;
    }
    case CSSM_OBS_BERNOULLI: {
      real p = (g > (real)6) ? (real)1 : (g < (real)-6) ? (real)0 : (real)1 / ((real)1 + r_exp<real>(-g));
      if (a.k0 != (real)0) return (p == (real)0) ? r_neg_big<real>() : r_log<real>(p);
      return (p == (real)1) ? r_neg_big<real>() : r_log<real>((real)1 - p);
    }
    default: return (real)0;
  }
}

// four N(0,1) per Philox block in fp32, two in fp64
template <typename real> struct Normals;
template <> struct Normals<float> {
  static constexpr int PER_CALL = 4;
  __device__ __forceinline__ static void draw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, float* z) {
    uint4 v = philox4x32_10(make_uint4(c0, c1, c2, c3), k0, k1);
    box_muller_f(v.x, v.y, z[0], z[1]);
    box_muller_f(v.z, v.w, z[2], z[3]);
  }
};
template <> struct Normals<double> {
  static constexpr int PER_CALL = 2;
  __device__ __forceinline__ static void draw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, double* z) {
    uint4 v = philox4x32_10(make_uint4(c0, c1, c2, c3), k0, k1);
    box_muller_d(v, z[0], z[1]);
  }
};

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
// K1  fused gather + propagate + f + log-weight + max
//     one thread = PPT consecutive particles (16-byte stores); the loads go through the ancestor
//     index (non-decreasing for systematic/stratified, so a warp still touches few sectors).
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
k_propagate_weight(const __grid_constant__ StepArgs<real> a, const real* __restrict__ xsrc, real* __restrict__ xdst,
                   const int32_t* __restrict__ anc, real* __restrict__ logw, const double* __restrict__ zinj,
                   long long N, long long Ns, unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step,
                   Scalars* __restrict__ sc) {
  constexpr int PPT = VecOf<real>::PPT;
  typedef typename VecOf<real>::type vec_t;
  constexpr int PC = Normals<real>::PER_CALL;
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PPT;
  const bool full = (i0 + PPT <= N);
  long long src[PPT];
  bool valid[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    valid[p] = (i0 + p < N);
    src[p] = i0 + p;
  }
  if (anc != nullptr) {
    if (full && PPT == 4) {
      int4 v = *reinterpret_cast<const int4*>(anc + i0);
      src[0] = v.x; src[1] = v.y; src[PPT > 2 ? 2 : 0] = v.z; src[PPT > 3 ? 3 : 0] = v.w;
    } else {
#pragma unroll
      for (int p = 0; p < PPT; ++p)
        if (valid[p]) src[p] = anc[i0 + p];
    }
  }
  real g[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) g[p] = (real)0;

  for (int kk = 0; kk < a.d; kk += 4) {
    // issue the loads of this chunk of (up to) 4 coordinates first
    real xv[4][PPT];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kk + j;
#pragma unroll
      for (int p = 0; p < PPT; ++p) xv[j][p] = (k < a.d && valid[p]) ? __ldg(xsrc + (long long)k * Ns + src[p]) : (real)0;
    }
    // noise: counter = (global slot, step, chunk)
    real z[4][PPT];
    if (zinj == nullptr) {
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        unsigned long long slot = slot0 + (unsigned long long)(i0 + p);
        real zz[4];
        if (PC == 4) {
          Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)(kk >> 2), key0, key1, zz);
        } else {
          Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)(kk >> 1), key0, key1, zz);
          if (kk + 2 < a.d)
            Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step, RNG_STEP | (uint32_t)((kk >> 1) + 1), key0, key1, zz + 2);
          else
            zz[2] = zz[3] = (real)0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j][p] = zz[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < PPT; ++p)
          z[j][p] = (kk + j < a.d && valid[p]) ? (real)zinj[(long long)(kk + j) * N + i0 + p] : (real)0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kk + j;
      if (k < a.d) {
        real A = a.A[k], M = a.M[k], D = a.D[k], S = a.S[k], Cc = a.C[k];
        real xn[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          real mean = r_fma<real>(A, xv[j][p] - M, M) + D;
          xn[p] = r_fma<real>(S, z[j][p], mean);
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
  if (!a.has_obs) return;

  real lw[PPT];
  double mx = -__longlong_as_double(0x7FF0000000000000ll);  // -inf
  bool bad = false;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    lw[p] = obs_loglik<real>(a, g[p]);
    if (valid[p]) {
      if (lw[p] != lw[p]) bad = true;
      else mx = fmax(mx, (double)lw[p]);
    }
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
  // block max -> one atomic per block
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  __shared__ double s_mx[8];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
  if (bad) s_bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m2 = s_mx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m2 = fmax(m2, s_mx[w]);
    atomicMax(&sc->gmax_key, ord_key(m2));
    if (s_bad) atomicOr(&sc->flags, FLAG_NAN_WEIGHT);
  }
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
k_lgcp_weight(const __grid_constant__ StepArgs<real> a, const real* __restrict__ xsrc, real* __restrict__ xdst,
              const int32_t* __restrict__ anc, real* __restrict__ logw, const double* __restrict__ zinj,
              const real* __restrict__ ctab, long long n_sub, real delta, long long N, long long Ns,
              unsigned long long slot0, uint32_t key0, uint32_t key1, uint32_t step, Scalars* __restrict__ sc) {
  constexpr int PC = Normals<real>::PER_CALL;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  double mx = -__longlong_as_double(0x7FF0000000000000ll);
  bool bad = false;
  if (valid) {
    long long src = anc ? (long long)anc[i] : i;
    real x[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) x[k] = (k < a.d) ? xsrc[(long long)k * Ns + src] : (real)0;
    unsigned long long slot = slot0 + (unsigned long long)i;
    real hz = (real)0;
    const uint32_t calls = (uint32_t)((a.d + PC - 1) / PC);
    for (long long s = 0; s < n_sub; ++s) {
      real gs = (real)0;
#pragma unroll
      for (int kk = 0; kk < DP; kk += PC) {
        real z[PC];
        if (kk < a.d) {
          if (zinj == nullptr) {
            Normals<real>::draw((uint32_t)slot, (uint32_t)(slot >> 32), step,
                                RNG_STEP | (uint32_t)(s * calls + (kk / PC)), key0, key1, z);
          } else {
#pragma unroll
            for (int j = 0; j < PC; ++j)
              z[j] = (kk + j < a.d) ? (real)zinj[((long long)s * a.d + kk + j) * N + i] : (real)0;
          }
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            int k = kk + j;
            if (k < DP && k < a.d) {
              real mean = r_fma<real>(a.A[k], x[k] - a.M[k], a.M[k]) + a.D[k];
              x[k] = r_fma<real>(a.S[k], z[j], mean);
              real cc = ctab ? ctab[s * a.d + k] : a.C[k];
              gs = r_fma<real>(cc, x[k], gs);
            }
          }
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
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  __shared__ double s_mx[8];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
  if (bad) s_bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m2 = s_mx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m2 = fmax(m2, s_mx[w]);
    atomicMax(&sc->gmax_key, ord_key(m2));
    if (s_bad) atomicOr(&sc->flags, FLAG_NAN_WEIGHT);
  }
}

// ---------------------------------------------------------------------------------------------
// weights: either w1 = exp_det(logw - gmax) (in-filter) or a caller-given fp64 array (cssm_resample)
// ---------------------------------------------------------------------------------------------
template <typename real>
struct WeightSrc {
  const real* logw;    // in-filter source (NULL when `direct` is used)
  const double* direct;
  double gmax;
  __device__ __forceinline__ double operator()(long long idx) const {
    if (direct) return direct[idx];
    return exp_det((double)logw[idx] - gmax);
  }
};

// max of a caller-given weight array -> sc->gmax_key (cssm_resample only)
__global__ void __launch_bounds__(256) k_max_direct(const double* __restrict__ w, long long N, Scalars* sc) {
  double mx = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) mx = fmax(mx, w[i]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  if ((threadIdx.x & 31) == 0) atomicMax(&sc->gmax_key, ord_key(mx));
}
// Before k_scan_tiles runs, the max lives in sc->gmax_key.  In-filter: gmax = that max, weights
// are <= 1, quantum 2^-96.  Direct weights: power-of-two pre-scale so that every weight is <= 1,
// qb = 96 - e with max = f * 2^e, f in [0.5, 1).
struct PreScan {
  double gmax;
  int qb;
};
__device__ __forceinline__ PreScan pre_scan(const Scalars* sc, bool direct) {
  double mx = ord_unkey(sc->gmax_key);
  PreScan p;
  p.gmax = direct ? 0.0 : mx;
  p.qb = 96;
  if (direct && mx > 1.0) {
    unsigned long long b = (unsigned long long)__double_as_longlong(mx);
    p.qb = 96 - ((int)((b >> 52) & 0x7ff) - 1022);
  }
  return p;
}

// K2a  total of the weights (exact).  Strided access: order is irrelevant for an exact sum.
template <typename real>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_total(const real* __restrict__ logw, const double* __restrict__ direct, long long N, Scalars* __restrict__ sc) {
  const PreScan ps = pre_scan(sc, direct != nullptr);
  WeightSrc<real> ws{logw, direct, ps.gmax};
  const int qb = ps.qb;
  const long long base = (long long)blockIdx.x * TILE;
  u128 acc = make_u128(0, 0);
#pragma unroll
  for (int j = 0; j < TILE_ITEMS; ++j) {
    long long idx = base + j * TILE_THREADS + threadIdx.x;
    if (idx < N) acc = add128(acc, fixq(ws(idx), qb));
  }
  acc = warp_sum128(acc);
  __shared__ u128 s_acc[TILE_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    u128 t = s_acc[0];
    for (int w = 1; w < TILE_THREADS / 32; ++w) t = add128(t, s_acc[w]);
    atomic_add128(&sc->tot, t);
  }
}

// K2b  per-tile sums of the (normalised) weights and the ESS accumulator
//      normalise = 1: CDF over wn = w/total (systematic, stratified; model/Resampling.scala:52-58)
//      normalise = 0: CDF over w itself       (multinomial; Breeze Multinomial uses raw params)
template <typename real>
__global__ void __launch_bounds__(TILE_THREADS)
k_tile_sums(const real* __restrict__ logw, const double* __restrict__ direct, long long N, int normalise,
            Scalars* __restrict__ sc, u128* __restrict__ tile_sum, double* __restrict__ tile_maxw) {
  const PreScan ps = pre_scan(sc, direct != nullptr);
  WeightSrc<real> ws{logw, direct, ps.gmax};
  const int qb = ps.qb;
  const double total = unfixq(sc->tot, qb);
  const long long base = (long long)blockIdx.x * TILE;
  u128 acc = make_u128(0, 0), acc2 = make_u128(0, 0);
  double mxw = 0.0;
#pragma unroll
  for (int j = 0; j < TILE_ITEMS; ++j) {
    long long idx = base + j * TILE_THREADS + threadIdx.x;
    if (idx < N) {
      double w = ws(idx);
      double wn = __ddiv_rn(w, total);
      acc = add128(acc, normalise ? fixq(wn, 96) : fixq(w, qb));
      acc2 = add128(acc2, fixq(__dmul_rn(wn, wn), 96));
      mxw = fmax(mxw, wn);
    }
  }
  acc = warp_sum128(acc);
  acc2 = warp_sum128(acc2);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mxw = fmax(mxw, __shfl_xor_sync(0xffffffffu, mxw, m));
  __shared__ u128 s_acc[TILE_THREADS / 32], s_acc2[TILE_THREADS / 32];
  __shared__ double s_mxw[TILE_THREADS / 32];
  if ((threadIdx.x & 31) == 0) {
    s_acc[threadIdx.x >> 5] = acc;
    s_acc2[threadIdx.x >> 5] = acc2;
    s_mxw[threadIdx.x >> 5] = mxw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u128 t = s_acc[0], t2 = s_acc2[0];
    double m2 = s_mxw[0];
    for (int w = 1; w < TILE_THREADS / 32; ++w) {
      t = add128(t, s_acc[w]);
      t2 = add128(t2, s_acc2[w]);
      m2 = fmax(m2, s_mxw[w]);
    }
    tile_sum[blockIdx.x] = t;
    tile_maxw[blockIdx.x] = m2;
    atomic_add128(&sc->ess_acc, t2);
  }
}

// K2c  one block: exclusive prefix over the tile sums, rounded tile-end CDF values, the ll
//      increment max + log(mean(w1)) and ESS = floor(1/sum wn^2)  (model/ParticleFilter.scala:127-128),
//      the resampling uniform, and the reset of the accumulators for the next step.
__global__ void __launch_bounds__(1024)
k_scan_tiles(Scalars* __restrict__ sc, const u128* __restrict__ tile_sum, u128* __restrict__ tile_excl,
             double* __restrict__ cend, int nt, long long N, int normalise, int direct, int add_ll, int use_u_inj,
             uint32_t key0, uint32_t key1, uint32_t step, double* __restrict__ ll_steps, int* __restrict__ ess_steps,
             long long step_slot) {
  __shared__ u128 s_part[1024];
  const int per = (nt + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(nt, b + per);
  u128 acc = make_u128(0, 0);
  for (int t = b; t < e; ++t) acc = add128(acc, tile_sum[t]);
  s_part[threadIdx.x] = acc;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 1024 partials (exact integers: order irrelevant)
  for (int off = 1; off < 1024; off <<= 1) {
    u128 v = make_u128(0, 0);
    if ((int)threadIdx.x >= off) v = s_part[threadIdx.x - off];
    __syncthreads();
    s_part[threadIdx.x] = add128(s_part[threadIdx.x], v);
    __syncthreads();
  }
  u128 run = (threadIdx.x == 0) ? make_u128(0, 0) : s_part[threadIdx.x - 1];
  const PreScan ps = pre_scan(sc, direct != 0);
  const int q = normalise ? 96 : ps.qb;
  for (int t = b; t < e; ++t) {
    tile_excl[t] = run;
    run = add128(run, tile_sum[t]);
    cend[t] = unfixq(run, q);
  }
  __syncthreads();  // every thread has read sc->gmax_key before thread 0 resets it
  if (threadIdx.x == 0) {
    const int qb = ps.qb;
    double total = unfixq(sc->tot, qb);
    double gmax = ps.gmax;
    double s2 = unfixq(sc->ess_acc, 96);
    sc->gmax = gmax;
    sc->qb = qb;
    double incr = gmax + log(total / (double)N);
    int flags = 0;
    if (!(total > 0.0) || gmax != gmax || gmax - gmax != 0.0) {  // all weights zero / NaN / infinite max
      incr = __longlong_as_double(0x7FF8000000000000ll);
      flags |= FLAG_ZERO_TOTAL;
    }
    sc->total = total;
    sc->ll_incr = incr;
    double inv = floor(1.0 / s2);
    int ess = (inv == inv && inv < 2147483647.0) ? (int)inv : (inv == inv ? 2147483647 : 0);  // Scala .toInt saturates, NaN -> 0
    if (add_ll) {
      sc->ll = sc->ll + incr;
      sc->ess = ess;
      if (ll_steps) ll_steps[step_slot] = sc->ll;
      if (ess_steps) ess_steps[step_slot] = ess;
    }
    if (flags) atomicOr(&sc->flags, flags);
    if (use_u_inj) {
      sc->u = sc->u_inj;
    } else {
      uint4 v = philox4x32_10(make_uint4(0u, 0u, step, RNG_RESAMPLE), key0, key1);
      sc->u = u64_to_unit_double(v.x, v.y);
    }
    // accumulators for the next step (gmax/total stay for K3)
    sc->gmax_key = 0ull;
    sc->tot = make_u128(0, 0);
    sc->ess_acc = make_u128(0, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// K3+K4  systematic / stratified: inclusive CDF of a tile in shared memory + ancestor search
// ---------------------------------------------------------------------------------------------
struct KFun {  // k_i of model/Resampling.scala:69 (systematic) and :82-83 (stratified)
  int kind;
  double u, n;
  const double* uarr;  // stratified, injected
  uint32_t key0, key1, step;
  __device__ __forceinline__ double ui(long long i) const {
    if (uarr) return uarr[i];
    uint4 v = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), step, RNG_RESAMPLE | 1u), key0, key1);
    return u64_to_unit_double(v.x, v.y);
  }
  __device__ __forceinline__ double operator()(long long i) const {
    if (kind == CSSM_RESAMPLE_SYSTEMATIC) return __ddiv_rn(__dadd_rn(u, (double)i), n);
    return __ddiv_rn(__dadd_rn((double)i, ui(i)), n);
  }
  // number of outputs i in [0, N) with k_i <= c   (k_i is non-decreasing in i)
  __device__ long long count_le(double c, long long N) const {
    if (!(c >= 0.0)) return 0;
    double est = c * n - (kind == CSSM_RESAMPLE_SYSTEMATIC ? u : 0.0);
    long long i = (est >= (double)N) ? N - 1 : (long long)floor(est);
    if (i < 0) i = 0;
    if (i > N - 1) i = N - 1;
    while (i + 1 < N && (*this)(i + 1) <= c) ++i;  // a few steps at most
    while (i >= 0 && (*this)(i) > c) --i;
    return i + 1;
  }
};

// inclusive CDF values (rounded fp64) of tile t into Cs[0..TILE); returns nothing, all threads call
template <typename real>
__device__ __forceinline__ void tile_cdf(const WeightSrc<real>& ws, int normalise, double total, int qb,
                                         const u128* __restrict__ tile_excl, int t, long long N, double* Cs,
                                         double* Ws, u128* s_warp) {
  const long long base = (long long)t * TILE + (long long)threadIdx.x * TILE_ITEMS;
  u128 e[TILE_ITEMS];
  u128 run = make_u128(0, 0);
#pragma unroll
  for (int j = 0; j < TILE_ITEMS; ++j) {
    long long idx = base + j;
    u128 q = make_u128(0, 0);
    double wv = 0.0;
    if (idx < N) {
      double w = ws(idx);
      wv = normalise ? __ddiv_rn(w, total) : w;
      q = normalise ? fixq(wv, 96) : fixq(w, qb);
    }
    if (Ws) Ws[threadIdx.x * TILE_ITEMS + j] = wv;
    run = add128(run, q);
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
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  u128 off = tile_excl[t];
  for (int w = 0; w < wid; ++w) off = add128(off, s_warp[w]);
  // exclusive within warp = inclusive - own total; recompute by adding the lower lanes instead
  u128 excl = shfl_up128(incl, 1);
  if (lane == 0) excl = make_u128(0, 0);
  off = add128(off, excl);
  const int q = normalise ? 96 : qb;
#pragma unroll
  for (int j = 0; j < TILE_ITEMS; ++j) Cs[threadIdx.x * TILE_ITEMS + j] = unfixq(add128(off, e[j]), q);
  __syncthreads();
}

// does adding w leave c unchanged?  This is what makes a TreeMap key repeat in the reference
// (model/Resampling.scala:55-57): the next cumulative sum equals the previous one.
__device__ __forceinline__ bool vanishes(double c, double w) { return __dadd_rn(c, w) == c; }

// mode 0: search (systematic / stratified), writes anc;  mode 1: write the CDF to cdf_out (multinomial)
template <typename real>
__global__ void __launch_bounds__(TILE_THREADS)
k_scan_search(const real* __restrict__ logw, const double* __restrict__ direct, long long N, int normalise,
              const Scalars* __restrict__ sc, const u128* __restrict__ tile_excl, const double* __restrict__ cend,
              const double* __restrict__ tile_maxw, int nt, int kind, const double* __restrict__ uarr, uint32_t key0,
              uint32_t key1, uint32_t step, int32_t* __restrict__ anc, double* __restrict__ cdf_out,
              int* __restrict__ flags_out) {
  __shared__ double Cs[TILE];
  __shared__ double Ws[TILE];
  __shared__ u128 s_warp[TILE_THREADS / 32];
  __shared__ long long s_lo, s_hi, s_pend, s_jfinal;
  __shared__ double s_wnext, s_c;
  __shared__ int s_tp, s_brk;

  const int t = blockIdx.x;
  WeightSrc<real> ws{logw, direct, sc->gmax};
  const int qb = sc->qb;
  const double total = sc->total;
  tile_cdf<real>(ws, normalise, total, qb, tile_excl, t, N, Cs, cdf_out ? nullptr : Ws, s_warp);
  const long long tile0 = (long long)t * TILE;
  const int tile_n = (int)min((long long)TILE, N - tile0);

  if (cdf_out != nullptr) {
    for (int j = threadIdx.x; j < tile_n; j += TILE_THREADS) cdf_out[tile0 + j] = Cs[j];
    return;
  }

  KFun kf{kind, sc->u, (double)N, uarr, key0, key1, step};
  const double c_end = Cs[tile_n - 1];
  if (threadIdx.x == 0) {
    s_lo = (t == 0) ? 0 : kf.count_le(cend[t - 1], N);
    s_hi = (t == nt - 1) ? N : kf.count_le(c_end, N);
    s_wnext = (t < nt - 1) ? __ddiv_rn(ws(tile0 + TILE), total) : 0.0;  // first weight of the next tile
    s_pend = 0x7FFFFFFFFFFFFFFFll;
  }
  __syncthreads();
  const long long lo = s_lo, hi = s_hi;
  // does the run of repeated keys at the end of this tile continue into the next tile?
  const bool cont = (t < nt - 1) && vanishes(c_end, s_wnext);
  if (t == nt - 1 && threadIdx.x == 0 && kf(N - 1) > c_end) atomicOr(flags_out, FLAG_CLAMPED);  // reference would throw (m.head)

  for (long long i = lo + threadIdx.x; i < hi; i += TILE_THREADS) {
    double k = kf(i);
    // first j with Cs[j] >= k (exists unless this is the clamped tail of the last tile)
    int a = 0, b = tile_n - 1;
    while (a < b) {
      int m = (a + b) >> 1;
      if (Cs[m] >= k) b = m; else a = m + 1;
    }
    // TreeMap: a duplicated key keeps the last particle inserted
    int j = a;
    while (j + 1 < tile_n && vanishes(Cs[j], Ws[j + 1])) ++j;
    if (cont && j == tile_n - 1) atomicMin(&s_pend, i);
    anc[i] = (int32_t)(tile0 + j);
  }
  __syncthreads();
  const long long pend = s_pend;
  if (pend >= hi) return;
  // The selected run of repeated keys continues past this tile; its last element is the ancestor.
  // Walk forward: whole tiles are skipped from the tables when every weight in them is strictly
  // below half an ulp of the running value (same binade), otherwise the tile is recomputed.
  if (threadIdx.x == 0) { s_tp = t + 1; s_c = c_end; s_jfinal = -1; }
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) {
      int tp = s_tp;
      double c = s_c;
      while (tp < nt) {
        double ce = cend[tp];
        // half ulp of c (c > 0 here; a zero running value never skips)
        long long cb = __double_as_longlong(c), eb = __double_as_longlong(ce);
        bool same_binade = (cb >> 52) == (eb >> 52) && ((cb >> 52) & 0x7ff) > 54;
        double half_ulp = same_binade ? __longlong_as_double((((cb >> 52) & 0x7ff) - 53) << 52) : 0.0;
        if (same_binade && tile_maxw[tp] < half_ulp) { c = ce; ++tp; } else break;
      }
      s_tp = tp;
      s_c = c;
      s_brk = TILE;
      if (tp >= nt) s_jfinal = N - 1;
    }
    __syncthreads();
    if (s_jfinal >= 0) break;
    const int tp = s_tp;
    const double c = s_c;
    tile_cdf<real>(ws, normalise, total, qb, tile_excl, tp, N, Cs, Ws, s_warp);
    const int tn = (int)min((long long)TILE, N - (long long)tp * TILE);
    // first element of tile tp that does NOT vanish against its predecessor's value
    for (int j = threadIdx.x; j < tn; j += TILE_THREADS) {
      double prev = (j == 0) ? c : Cs[j - 1];
      if (!vanishes(prev, Ws[j])) atomicMin(&s_brk, j);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_brk < tn) s_jfinal = (long long)tp * TILE + s_brk - 1;
      else if (tp == nt - 1) s_jfinal = N - 1;
      else { s_c = Cs[tn - 1]; s_tp = tp + 1; }
    }
    __syncthreads();
    if (s_jfinal >= 0) break;
  }
  const long long jfinal = s_jfinal;
  for (long long i = pend + threadIdx.x; i < hi; i += TILE_THREADS) anc[i] = (int32_t)jfinal;
}

// K4'  multinomial: Breeze Multinomial.draw first-draw walk = first j with cumulative >= u*sum
__global__ void __launch_bounds__(256)
k_multinomial_search(const double* __restrict__ cdf, long long N, const double* __restrict__ uarr, uint32_t key0,
                     uint32_t key1, uint32_t step, int32_t* __restrict__ anc, int* __restrict__ flags_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double u;
  if (uarr) u = uarr[i];
  else {
    uint4 v = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), step, RNG_RESAMPLE | 1u), key0, key1);
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
// K5  gather (only when the resampled cloud has to be materialised: get_particles, shard export)
// ---------------------------------------------------------------------------------------------
template <typename real, typename out_t>
__global__ void __launch_bounds__(256)
k_gather(const real* __restrict__ x, const int32_t* __restrict__ anc, out_t* __restrict__ out, int d, long long N,
         long long Ns, long long out_stride) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  long long s = anc ? (long long)anc[i] : i;
  for (int k = 0; k < d; ++k) out[(long long)k * out_stride + i] = (out_t)x[(long long)k * Ns + s];
}

// copy with dtype conversion (logw / propagated state read-back)
template <typename real>
__global__ void __launch_bounds__(256) k_to_double(const real* __restrict__ in, double* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}
template <typename real>
__global__ void __launch_bounds__(256)
k_w1_out(const real* __restrict__ logw, const Scalars* __restrict__ sc, double* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = exp_det((double)logw[i] - sc->gmax);
}

// Resampling.sampleOne (model/Resampling.scala:151-154): one uniformly chosen particle of the
// current (resampled) cloud -> out[d] (double)
template <typename real>
__global__ void k_sample_one(const real* __restrict__ x, const int32_t* __restrict__ anc, double* __restrict__ out, int d,
                             long long N, long long Ns, uint32_t key0, uint32_t key1, uint32_t step) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint4 v = philox4x32_10(make_uint4(0u, 0u, step, RNG_SAMPLE_ONE), key0, key1);
  unsigned long long r = ((unsigned long long)v.x << 32) | v.y;
  long long i = (long long)(r % (unsigned long long)N);
  long long s = anc ? (long long)anc[i] : i;
  for (int k = 0; k < d; ++k) out[k] = (double)x[(long long)k * Ns + s];
}

// per-coordinate mean of the resampled cloud (ParticleFilter.meanState); fp64 accumulation
template <typename real>
__global__ void __launch_bounds__(256)
k_mean_state(const real* __restrict__ x, const int32_t* __restrict__ anc, double* __restrict__ out, int d, long long N,
             long long Ns) {
  const int k = blockIdx.y;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    long long s = anc ? (long long)anc[i] : i;
    acc += (double)x[(long long)k * Ns + s];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + k, acc / (double)N);
}

}  // namespace cssm
