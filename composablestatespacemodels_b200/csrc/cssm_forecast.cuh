// cssm_forecast.cuh -- the forecast step that follows the filter in the streaming examples:
//
//   ParticleFilter.getForecast      model/ParticleFilter.scala:368-383   per particle: x1 = stepFunction(dt)(x).draw,
//                                                                        gamma = f(x1, t), eta = link(gamma),
//                                                                        obs ~ observation(gamma)
//   ParticleFilter.getMeanForecast  :394-412                             mean / credible intervals of x1, eta and of
//                                                                        a SECOND observation draw per particle
//   SimulateData.simStep / forecast model/Data.scala:186-217             the same step chained over several times
//
// One kernel, k_forecast, writes a forecast cloud of d + 4 columns [x1 (d) | gamma | eta | obs | obs2] (SoA, filter
// dtype) next to the filter's own cloud, which it does not touch; the summaries are the mean and radix-select
// kernels of cssm_kernels.cuh run on that cloud.  The observation samplers follow the distributions the
// reference draws from (model/Model.scala: Poisson :267, Gamma-Poisson mixture :169-178, Gaussian :209-213,
// :242-246, Bernoulli :316, scaled Student's t :145-149, zero-inflated Poisson :282-290, Beta :340-341); the
// reference's own RNG streams (Breeze / JDK) are not reproduced, a forecast is compared in distribution.
#pragma once
#include "cssm_kernels.cuh"

namespace cssm {

enum : uint32_t { RNG_FORECAST = 5u << 24 };

// sequential draws of one particle: Philox counter = (slot, step, purpose | call number)
struct RngStream {
  uint32_t c0, c1, step, k0, k1, ncall;
  uint4 buf;
  int pos;
  double spare;
  bool has_spare;
  __device__ __forceinline__ void init(unsigned long long slot, uint32_t step_, uint32_t key0, uint32_t key1) {
    c0 = (uint32_t)slot; c1 = (uint32_t)(slot >> 32); step = step_; k0 = key0; k1 = key1;
    ncall = 0u; pos = 4; has_spare = false; spare = 0.0;
    buf = make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ uint32_t u32() {
    if (pos == 4) {
      buf = philox4x32(make_uint4(c0, c1, step, RNG_FORECAST | (ncall & 0xFFFFFFu)), k0, k1);
      ++ncall;
      pos = 0;
    }
    const uint32_t v = (pos == 0) ? buf.x : (pos == 1) ? buf.y : (pos == 2) ? buf.z : buf.w;
    ++pos;
    return v;
  }
  // (0, 1), 53 bits
  __device__ __forceinline__ double uniform() {
    const unsigned long long a = u32(), b = u32();
    return ((double)(((a << 32) | b) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ __forceinline__ double normal() {
    if (has_spare) { has_spare = false; return spare; }
    const double u1 = uniform(), u2 = uniform();
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    spare = r * s;
    has_spare = true;
    return r * c;
  }
};

// Poisson(lam): inversion of the CDF below 10, Hoermann's transformed rejection (PTRS, 1993) above
__device__ __noinline__ double sample_poisson(RngStream& g, double lam) {
  if (!(lam >= 0.0) || lam > 1e300) return __longlong_as_double(0x7FF8000000000000ll);  // Breeze requires a finite mean >= 0
  if (lam == 0.0) return 0.0;
  if (lam < 10.0) {
    const double u = g.uniform();
    double p = exp(-lam), F = p;
    int k = 0;
    while (u > F && k < 200) { ++k; p *= lam / (double)k; F += p; }
    return (double)k;
  }
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 256; ++it) {
    const double U = g.uniform() - 0.5, V = g.uniform();
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
  }
  return floor(lam);  // not reached in practice (acceptance > 0.86 per round)
}

// Gamma(shape, 1): Marsaglia & Tsang (2000); shape < 1 through Gamma(shape + 1) * U^(1/shape)
__device__ __noinline__ double sample_gamma(RngStream& g, double shape) {
  if (!(shape > 0.0)) return __longlong_as_double(0x7FF8000000000000ll);
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(g.uniform(), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 256; ++it) {
    const double x = g.normal();
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = g.uniform();
    if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d;
}

// Model.link (model/Model.scala:23,182,269,294,318-326,345)
__device__ __forceinline__ double link_of(int obs_kind, double gamma) {
  switch (obs_kind) {
    case CSSM_OBS_POISSON: case CSSM_OBS_NEGBIN: case CSSM_OBS_ZIP: return exp(gamma);
    case CSSM_OBS_BERNOULLI: return (gamma > 6.0) ? 1.0 : (gamma < -6.0) ? 0.0 : 1.0 / (1.0 + exp(-gamma));
    case CSSM_OBS_BETA: return exp(-gamma);
    default: return gamma;
  }
}

struct ObsDraw {
  int obs_kind, obs_df, has_scale;
  double scale;  // raw, as in the descriptor
};

// one draw of Model.observation(gamma)
__device__ __forceinline__ double sample_observation(RngStream& g, const ObsDraw& o, double gamma) {
  switch (o.obs_kind) {
    case CSSM_OBS_POISSON: return sample_poisson(g, exp(gamma));
    case CSSM_OBS_NEGBIN: {  // lambda ~ Gamma(size, prob / (1 - prob)) = Gamma(size, mu / size), then Poisson(lambda)
      const double size = exp(o.scale), mu = exp(gamma);
      return sample_poisson(g, sample_gamma(g, size) * (mu / size));
    }
    case CSSM_OBS_NORMAL: return gamma + exp(o.scale) * g.normal();
    case CSSM_OBS_BERNOULLI: return (g.uniform() < link_of(o.obs_kind, gamma)) ? 1.0 : 0.0;
    case CSSM_OBS_STUDENT_T: {  // StudentsT(df) * v + gamma; t = z / sqrt(chi2_df / df), chi2_df = 2 Gamma(df / 2)
      const double df = (double)o.obs_df;
      const double z = g.normal();
      const double chi2 = 2.0 * sample_gamma(g, 0.5 * df);
      return z / sqrt(chi2 / df) * exp(o.scale) + gamma;
    }
    case CSSM_OBS_ZIP: {
      const double ev = exp(o.scale), p = ev / (1.0 + ev);
      const double u = g.uniform();
      const double nz = sample_poisson(g, exp(gamma));
      return (u < p) ? 0.0 : nz;
    }
    case CSSM_OBS_BETA: {  // Beta(link(gamma), beta) with beta = the raw scale
      const double ga = sample_gamma(g, exp(-gamma)), gb = sample_gamma(g, o.scale);
      return ga / (ga + gb);
    }
    default: return __longlong_as_double(0x7FF8000000000000ll);  // LGCP: `observation = ???` in the reference
  }
}

// src_fc == NULL: start from the filter's current cloud (through the ancestors of the last resampling);
// otherwise continue from the forecast cloud itself (Data.forecast's scan), in place.
// a.A/D/S: transition over dt, a.C: f-coefficients at the forecast time.
template <typename real>
__global__ void __launch_bounds__(256)
k_forecast(const __grid_constant__ StepArgs<real> a, const __grid_constant__ Peers pr, const int32_t* __restrict__ anc,
           real* __restrict__ fc, int from_fc, ObsDraw od, long long N, long long Ns, uint32_t key0, uint32_t key1,
           uint32_t step) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int d = a.d;
  const real* src;
  long long sstride = Ns;
  if (from_fc) {
    src = fc + i;
  } else {
    src = reinterpret_cast<const real*>(pr.x[pr.rank]) + i;
    if (anc) src = reinterpret_cast<const real*>(pr.x[0]) + anc[i];
  }
  RngStream g;
  g.init((unsigned long long)i, step, key0, key1);
  real gam = (real)0;
  for (int k = 0; k < d; ++k) {
    const real z = (real)g.normal();
    const real xn = r_fma<real>(a.S[k], z, r_fma<real>(a.A[k], src[(long long)k * sstride], a.D[k]));
    gam = r_fma<real>(a.C[k], xn, gam);
    fc[(long long)k * Ns + i] = xn;
  }
  const double gd = (double)gam;
  fc[(long long)d * Ns + i] = gam;
  fc[(long long)(d + 1) * Ns + i] = (real)link_of(od.obs_kind, gd);
  fc[(long long)(d + 2) * Ns + i] = (real)sample_observation(g, od, gd);
  fc[(long long)(d + 3) * Ns + i] = (real)sample_observation(g, od, gd);
}

// ---------------------------------------------------------------------------------------------
// FilterInterpolate (model/ParticleFilter.scala:273-311): every particle is a PATH (List[State], newest first) and
// an observed step resamples whole paths.  The device keeps the propagated cloud of every step, px[s] (s = 0 is
// the initial cloud), and the ancestors of every resampling, panc[s-1]; a path is read back by walking the
// ancestor tree from the newest step to the oldest.  out[p][s][k], s = 0 .. len (oldest state first).
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
k_paths(const real* __restrict__ px, const int32_t* __restrict__ panc, const uint8_t* __restrict__ resampled, int len, long long N,
        long long Ns, int d, const int32_t* __restrict__ idx, long long n_idx, double* __restrict__ out) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_idx) return;
  long long j = idx ? (long long)idx[p] : p;
  const size_t cloud = (size_t)d * (size_t)Ns;
  for (int s = len; s >= 0; --s) {
    if (s > 0 && resampled[s - 1]) j = panc[(size_t)(s - 1) * (size_t)N + j];
    const real* x = px + (size_t)s * cloud + j;
    double* o = out + ((size_t)p * (size_t)(len + 1) + (size_t)s) * (size_t)d;
    for (int k = 0; k < d; ++k) o[k] = (double)x[(size_t)k * Ns];
  }
}

}  // namespace cssm
