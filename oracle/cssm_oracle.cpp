// cssm_oracle.cpp -- CPU ORACLE for the particle-filter hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a from-scratch, single-threaded restatement of the reference algorithm
// (jonnylaw/ComposableStateSpaceModels, 100 % Scala) used as the parity checker for the CUDA
// library and as the CPU baseline of bench.py.  Nothing in the product path may call it: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// PARITY UNPINNED.  The reference cannot be compiled or run here (no JVM, no dependency jars),
// and its own tests pin no numeric result of this path (src/test/scala/SamplingTest.scala:12-22
// only checks that resampling preserves the vector length).  The restatement is therefore
// anchored on (i) the reference source, cited function by function below, (ii) the published
// definitions of the third-party arithmetic it calls -- org.scalanlp:breeze_2.13:1.0
// (build.sbt:40): Poisson.logProbabilityOf, Gaussian.logPdf, Multinomial.draw, lgamma -- which
// tests/ cross-check against scipy, and (iii) an exact Kalman-filter likelihood for the
// linear-Gaussian compositions.  Citations: "model/X.scala" =
// src/main/scala/com/github/jonnylaw/model/X.scala of the reference.
//
// Two summation orders are provided for everything that feeds the ancestor search:
//   ORC_ORDER_REFERENCE  the reference's sequential fp64 foldLeft/scanLeft, literally;
//   ORC_ORDER_DEVICE     the order-invariant definition the GPU uses.  The weights
//                        w1 = exp(logw - max) are summed EXACTLY as 2^-96 fixed-point integers
//                        (any association -- any tile size, grid size or GPU count -- gives the
//                        same integer); a cumulative value is that exact integer S_j converted by
//                        dbl128 (below; within 1.5 ulp, monotone), P_j = dbl128(S_j), and
//                        total = P_{N-1}.  The reference normalises first and compares
//                        C_j = cumsum(w/total)_j >= k_i; the device compares the same inequality
//                        multiplied through by the total, P_j >= fl(k_i * total), which saves a
//                        division per particle and differs from the reference only when a key is
//                        within a few ulps of a cumulative value.  ESS = floor(1/(sum w^2 /
//                        total^2)) with sum w^2 exact as well.  Here all of it is a plain
//                        sequential loop over unsigned __int128.  The TreeMap "duplicate key"
//                        rule is applied through the test that creates duplicate keys in the
//                        reference, fl(C_j + wn_{j+1}) == C_j, evaluated in the reference's
//                        normalised domain: C_j = fl(P_j/total), wn = fl(w/total).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <algorithm>
#include <thread>

#include "../include/cssm.h"

typedef unsigned __int128 u128;

#define ORC_ORDER_REFERENCE 0
#define ORC_ORDER_DEVICE 1      // device definition, fp64 weights (F64 filters, cssm_resample)
#define ORC_ORDER_DEVICE_F32 2  // device definition, weights w1 evaluated in fp32 (F32 filters)

extern "C" {

// ---------------------------------------------------------------------------------------------
// Deterministic exp for x <= 0 (DEVICE order only).  Uses nothing but IEEE-754 fma/mul/add and
// bit assembly, so the CUDA library evaluates the identical sequence and gets identical bits.
// |error| < 1 ulp; the reference calls breeze.numerics.exp = java.lang.Math.exp (<= 1 ulp).
// ---------------------------------------------------------------------------------------------
double orc_exp_det(double x) {
  if (x != x) return x;
  if (x < -745.5) return 0.0;
  const double LOG2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  double kf = std::nearbyint(x * LOG2E);
  double r = std::fma(kf, -LN2_HI, x);
  r = std::fma(kf, -LN2_LO, r);
  // Taylor to degree 13 on |r| <= ln2/2: truncation 4e-18 relative
  double p = 1.0 / 6227020800.0;
  p = std::fma(p, r, 1.0 / 479001600.0);
  p = std::fma(p, r, 1.0 / 39916800.0);
  p = std::fma(p, r, 1.0 / 3628800.0);
  p = std::fma(p, r, 1.0 / 362880.0);
  p = std::fma(p, r, 1.0 / 40320.0);
  p = std::fma(p, r, 1.0 / 5040.0);
  p = std::fma(p, r, 1.0 / 720.0);
  p = std::fma(p, r, 1.0 / 120.0);
  p = std::fma(p, r, 1.0 / 24.0);
  p = std::fma(p, r, 1.0 / 6.0);
  p = std::fma(p, r, 0.5);
  p = std::fma(p, r, 1.0);
  p = std::fma(p, r, 1.0);
  int k = (int)kf;
  // scale by 2^k in two exact steps so that subnormal results round once
  int k1 = k / 2, k2 = k - k1;
  uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
  double s1, s2;
  std::memcpy(&s1, &b1, 8);
  std::memcpy(&s2, &b2, 8);
  return (p * s1) * s2;
}

// The F32 filter evaluates w1 = exp(logw - max) in fp32 (its log-weights are fp32 anyway): same
// construction with fmaf only; arguments below -86 give exactly 0.  Everything after w1 (exact
// sums, CDF, keys) is the fp64 / integer arithmetic of ORC_ORDER_DEVICE applied to (double)w1.
float orc_expf_det(float x) {
  if (!(x >= -86.0f)) return (x != x) ? x : 0.0f;
  const float LOG2E = 1.442695040888963f, NLN2_HI = -0.693145751953125f, NLN2_LO = -1.428606765330187e-06f;
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23: adding it rounds to the nearest integer (ties to even)
  volatile float prod = x * LOG2E;
  volatile float t = prod + MAGIC;
  float kf = t - MAGIC;
  float r = std::fmaf(kf, NLN2_HI, x);
  r = std::fmaf(kf, NLN2_LO, r);
  float p = 1.0f / 5040.0f;
  p = std::fmaf(p, r, 1.0f / 720.0f);
  p = std::fmaf(p, r, 1.0f / 120.0f);
  p = std::fmaf(p, r, 1.0f / 24.0f);
  p = std::fmaf(p, r, 1.0f / 6.0f);
  p = std::fmaf(p, r, 0.5f);
  p = std::fmaf(p, r, 1.0f);
  p = std::fmaf(p, r, 1.0f);
  int k = (int)kf;  // in [-124, 0]: the scaled result is a normal fp32 number
  uint32_t b;
  std::memcpy(&b, &p, 4);
  b += (uint32_t)k << 23;
  std::memcpy(&p, &b, 4);
  return p;
}

// floor(x * 2^q) for 0 <= x, result must fit 128 bits (callers guarantee x*2^q < 2^100)
static inline u128 fixq(double x, int q) {
  if (!(x > 0.0)) return 0;  // 0, negative, NaN
  uint64_t b;
  std::memcpy(&b, &x, 8);
  int ef = (int)((b >> 52) & 0x7ff);
  if (ef == 0) return 0;  // subnormal: below any quantum we use
  uint64_t mant = (b & 0xFFFFFFFFFFFFFull) | (1ull << 52);
  int sh = ef - 1075 + q;  // x = mant * 2^(ef-1075)
  if (sh >= 0) return (u128)mant << sh;
  if (sh <= -53) return 0;
  return (u128)(mant >> (-sh));
}

// round-to-nearest-even of e * 2^-q
static inline double unfixq(u128 e, int q) {
  if (e == 0) return 0.0;
  uint64_t hi = (uint64_t)(e >> 64), lo = (uint64_t)e;
  int p = hi ? 127 - __builtin_clzll(hi) : 63 - __builtin_clzll(lo);  // msb position
  uint64_t mant;
  int s = 0;
  if (p <= 52) {
    mant = lo;
  } else {
    s = p - 52;
    mant = (uint64_t)(e >> s);
    u128 rem = e & (((u128)1 << s) - 1), half = (u128)1 << (s - 1);
    if (rem > half || (rem == half && (mant & 1))) mant += 1;
  }
  return std::ldexp((double)mant, s - q);
}

// the device's conversion of an exact 128-bit sum to fp64: (double)hi * 2^64 + (double)lo, then
// the exact power-of-two scale.  Three correctly rounded IEEE operations (two conversions, one
// add), so CPU and GPU agree bit for bit; monotone in e; within 1.5 ulp of e * 2^-q.
static inline double dbl128(u128 e, int q) {
  uint64_t hi = (uint64_t)(e >> 64), lo = (uint64_t)e;
  volatile double a = (double)hi * 18446744073709551616.0;
  volatile double b = (double)lo;
  volatile double c = a + b;
  return std::ldexp(c, -q);
}
double orc_dbl128(uint64_t lo, uint64_t hi, int q) { return dbl128(((u128)hi << 64) | lo, q); }

void orc_fix96(double x, uint64_t* lo, uint64_t* hi) {
  u128 e = fixq(x, 96);
  *lo = (uint64_t)e;
  *hi = (uint64_t)(e >> 64);
}
double orc_unfix96(uint64_t lo, uint64_t hi) { return unfixq(((u128)hi << 64) | lo, 96); }

// ---------------------------------------------------------------------------------------------
// model helpers
// ---------------------------------------------------------------------------------------------
int orc_dim(const cssm_model_desc_t* m) {
  int d = 0;
  for (int l = 0; l < m->n_leaves; ++l) d += m->leaves[l].dim;
  return d;
}

// a2  Sde.initialState: BM model/Sde.scala:104-108, OU :152-156, GenBM :75-80, composed :206-209
//     x0 = sqrt(c0) * z + m0, leaves left to right.   x, z0: [d][N]
void orc_init_state(const cssm_model_desc_t* m, int64_t N, const double* z0, double* x) {
  int k = 0;
  for (int l = 0; l < m->n_leaves; ++l) {
    const cssm_leaf_t& L = m->leaves[l];
    for (int c = 0; c < L.dim; ++c, ++k)
      for (int64_t i = 0; i < N; ++i) x[k * N + i] = std::sqrt(L.c0[c]) * z0[k * N + i] + L.m0[c];
  }
}

// a4  exact transitions, per particle and per coordinate exactly as written in the reference
//     (variance(dt) is re-evaluated per particle there; the value is the same).
static inline double step_exact_1(const cssm_leaf_t& L, int c, double dt, double x, double z) {
  switch (L.sde_kind) {
    case CSSM_SDE_BROWNIAN: {  // model/Sde.scala:114-123
      double sd = std::sqrt(L.sigma[c] * dt);
      return sd * z + x;
    }
    case CSSM_SDE_GEN_BROWNIAN: {  // :86-95
      double mean = x + L.mu[c] * dt;
      double sd = std::sqrt(L.sigma[c] * dt);
      return sd * z + mean;
    }
    default: {  // OU :139-150
      double phi = L.phi[c], mu = L.mu[c], sigma = L.sigma[c];
      double var = (sigma * sigma / (phi * 2.0)) * (1.0 - std::exp(phi * -2.0 * dt));
      double mean = mu + (x - mu) * std::exp(-phi * dt);
      return std::sqrt(var) * z + mean;
    }
  }
}

// a5  Euler-Maruyama, model/Sde.scala:30-43 with drift/diffusion of :82-84,:110-112,:158-162
//     (Brownian drift is the constant 1.0 there, sic; sigma is used as a standard deviation).
static inline double step_euler_1(const cssm_leaf_t& L, int c, double dt, double x, double z) {
  double drift;
  switch (L.sde_kind) {
    case CSSM_SDE_BROWNIAN: drift = 1.0; break;
    case CSSM_SDE_GEN_BROWNIAN: drift = L.mu[c]; break;
    default: drift = L.phi[c] * (L.mu[c] - x); break;
  }
  double wiener = std::sqrt(dt) * z;  // dW, :30-33 / :233-238
  double a = drift * dt;
  double b = L.sigma[c] * wiener;
  return (x + a) + b;
}

// composed stepFunction, model/Sde.scala:223-229 (left subtree then right).  x_in, z, x_out [d][N]
void orc_propagate(const cssm_model_desc_t* m, int64_t N, double dt, const double* x_in,
                   const double* z, double* x_out) {
  int k = 0;
  for (int l = 0; l < m->n_leaves; ++l) {
    const cssm_leaf_t& L = m->leaves[l];
    for (int c = 0; c < L.dim; ++c, ++k)
      for (int64_t i = 0; i < N; ++i)
        x_out[k * N + i] = (m->step_mode == CSSM_STEP_EULER)
                               ? step_euler_1(L, c, dt, x_in[k * N + i], z[k * N + i])
                               : step_exact_1(L, c, dt, x_in[k * N + i], z[k * N + i]);
  }
}

// a6  Model.f: composed sum over leaves (model/Model.scala:122-128), first component
//     (:184,250,271,328,366) or seasonal buildF(h,t) dot x (:217-225)
static inline double f_one(const cssm_model_desc_t* m, int64_t N, const double* x, int64_t i, double t) {
  double g = 0.0;
  int k = 0;
  for (int l = 0; l < m->n_leaves; ++l) {
    const cssm_leaf_t& L = m->leaves[l];
    double fl;
    if (L.f_kind == CSSM_F_SEASONAL) {
      double frequency = 2 * M_PI / L.period;
      fl = 0.0;
      for (int a = 1; a <= L.harmonics; ++a) {
        fl += std::cos(frequency * a * t) * x[(k + 2 * (a - 1)) * N + i];
        fl += std::sin(frequency * a * t) * x[(k + 2 * (a - 1) + 1) * N + i];
      }
    } else {
      fl = x[k * N + i];
    }
    g = (l == 0) ? fl : g + fl;
    k += L.dim;
  }
  return g;
}
void orc_f(const cssm_model_desc_t* m, int64_t N, const double* x, double t, double* gamma) {
  for (int64_t i = 0; i < N; ++i) gamma[i] = f_one(m, N, x, i, t);
}

// a7  dataLikelihood of the left-most model (model/Model.scala:118-132)
double orc_loglik_1(const cssm_model_desc_t* m, double g, double y) {
  switch (m->obs_kind) {
    case CSSM_OBS_POISSON: {  // :269-273, Breeze Poisson.logProbabilityOf(k) = -mean + k*log(mean) - lgamma(k+1)
      int k = (int)y;
      double mean = std::exp(g);
      return -mean + k * std::log(mean) - std::lgamma(k + 1.0);
    }
    case CSSM_OBS_NEGBIN: {  // :186-195
      int k = (int)y;
      double size = std::exp(m->scale);
      double mu = std::exp(g);
      return std::lgamma(size + k) - std::lgamma(k + 1.0) - std::lgamma(size) +
             size * std::log(size / (mu + size)) + k * std::log(mu / (mu + size));
    }
    case CSSM_OBS_NORMAL: {  // :227-233, :252-258; Breeze Gaussian(mu, sigma).logPdf
      double v = std::exp(m->scale);
      double d = (y - g) / v;
      return -d * d / 2.0 - (std::log(std::sqrt(2 * M_PI)) + std::log(v));
    }
    case CSSM_OBS_BERNOULLI: {  // :318-336
      double p = (g > 6) ? 1.0 : (g < -6) ? 0.0 : 1.0 / (1 + std::exp(-g));
      if (y == 1.0) return (p == 0.0) ? -1e99 : std::log(p);
      return (p == 1.0) ? -1e99 : std::log(1 - p);
    }
    case CSSM_OBS_STUDENT_T: {  // :154-160: 1/v * StudentsT(df).logPdf((y - eta)/v)  (the 1/v factor is the reference's)
      // Breeze StudentsT(df).logPdf(x) = lgamma((df+1)/2) - lgamma(df/2) - log(pi df)/2 - (df+1)/2 * log(1 + x^2/df)
      double v = std::exp(m->scale), df = (double)m->obs_df;
      double x = (y - g) / v;
      double lp = std::lgamma((df + 1.0) / 2.0) - std::lgamma(df / 2.0) - 0.5 * std::log(M_PI * df) -
                  (df + 1.0) / 2.0 * std::log(1.0 + x * x / df);
      return 1.0 / v * lp;
    }
    case CSSM_OBS_ZIP: {  // :298-306
      double p = std::exp(m->scale) / (1 + std::exp(m->scale));
      int k = (int)y;
      if (k == 0) return std::log(p + (1 - p) * std::exp(-std::exp(g)));
      return -std::log(1 + std::exp(m->scale)) + k * g - std::exp(g) - std::lgamma(k + 1.0);
    }
    case CSSM_OBS_BETA: {  // :349-352: new Beta(exp(-gamma), 1.0).logPdf(y);
      // Breeze Beta(a, b).logPdf(x) = (a-1) log x + (b-1) log(1-x) - (lgamma(a) + lgamma(b) - lgamma(a+b))
      double a = std::exp(-g), b = 1.0;
      return (a - 1) * std::log(y) + (b - 1) * std::log(1 - y) - (std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b));
    }
    default: return 0.0 / 0.0;  // LGCP has no dataLikelihood (:368)
  }
}
void orc_loglik(const cssm_model_desc_t* m, int64_t N, const double* gamma, double y, double* logw) {
  for (int64_t i = 0; i < N; ++i) logw[i] = orc_loglik_1(m, gamma[i], y);
}

// Model.link (model/Model.scala:25, and the overrides :180,:268,:296,:318-322,:345)
double orc_link(const cssm_model_desc_t* m, double g) {
  switch (m->obs_kind) {
    case CSSM_OBS_POISSON: case CSSM_OBS_NEGBIN: case CSSM_OBS_ZIP: return std::exp(g);
    case CSSM_OBS_BERNOULLI: return (g > 6) ? 1.0 : (g < -6) ? 0.0 : 1.0 / (1 + std::exp(-g));
    case CSSM_OBS_BETA: return std::exp(-g);
    default: return g;
  }
}

// ParticleFilter.getIntervals (model/ParticleFilter.scala:415-424) of a cloud x[d][N] at time t:
// meanState (:478-480, weightedMean with unit weights :465-473), getallCredibleIntervals (:490-513:
// per coordinate sorted(n - index - 1), sorted(index - 1), index = floor(interval * n)),
// meanEta = link(f(stateMean, t)), getOrderStatistic of eta_i = link(f(x_i, t)) (:455-460:
// ordered(n - index), ordered(index), index = floor(n * interval)).  eta[3] = {meanEta, lower, upper}.
// Returns -1 where the reference would throw IndexOutOfBounds.
int orc_intervals(const cssm_model_desc_t* m, int64_t N, const double* x, double t, double interval, double* mean,
                  double* lower, double* upper, double* eta) {
  const int d = orc_dim(m);
  const int64_t idx_s = (int64_t)std::floor(interval * (double)N), idx_e = (int64_t)std::floor((double)N * interval);
  if (N - idx_s - 1 < 0 || N - idx_s - 1 >= N || idx_s - 1 < 0 || idx_s - 1 >= N || N - idx_e < 0 || N - idx_e >= N || idx_e >= N)
    return -1;
  double wsum = 0.0;
  for (int64_t i = 0; i < N; ++i) wsum = wsum + 1.0;
  const double wn = 1.0 / wsum;
  std::vector<double> col(N);
  for (int k = 0; k < d; ++k) {
    double acc = x[(int64_t)k * N] * wn;
    for (int64_t i = 1; i < N; ++i) acc = acc + x[(int64_t)k * N + i] * wn;
    mean[k] = acc;
    for (int64_t i = 0; i < N; ++i) col[i] = x[(int64_t)k * N + i];
    std::sort(col.begin(), col.end());
    lower[k] = col[N - idx_s - 1];
    upper[k] = col[idx_s - 1];
  }
  for (int64_t i = 0; i < N; ++i) col[i] = orc_link(m, f_one(m, N, x, i, t));
  std::sort(col.begin(), col.end());
  eta[0] = orc_link(m, f_one(m, 1, mean, 0, t));
  eta[1] = col[N - idx_e];
  eta[2] = col[idx_e];
  return 0;
}

// a8  max, w1 = exp(w - max)  (model/ParticleFilter.scala:124-125)
double orc_max(int64_t N, const double* w) {
  double mx = w[0];
  for (int64_t i = 1; i < N; ++i)
    if (w[i] > mx) mx = w[i];
  return mx;
}
void orc_w1(int64_t N, const double* logw, double mx, int order, double* w1) {
  for (int64_t i = 0; i < N; ++i) {
    if (order == ORC_ORDER_DEVICE) w1[i] = orc_exp_det(logw[i] - mx);
    else if (order == ORC_ORDER_DEVICE_F32) {
      volatile float x = (float)logw[i] - (float)mx;  // log-weights of an F32 filter are fp32 values
      w1[i] = (double)orc_expf_det(x);
    } else w1[i] = std::exp(logw[i] - mx);
  }
}

// power-of-two pre-scale so that every weight is <= 1 before it is made fixed point
static int weight_shift(int64_t N, const double* w) {
  double mx = 0.0;
  for (int64_t i = 0; i < N; ++i)
    if (w[i] > mx) mx = w[i];
  if (!(mx > 1.0)) return 0;
  int e;
  std::frexp(mx, &e);  // mx = f * 2^e, f in [0.5,1)  ->  mx * 2^-e < 1
  return e;
}

// total of the weights: Resampling.normalise's foldLeft (model/Resampling.scala:22) / Seq.sum
double orc_total(int64_t N, const double* w, int order) {
  if (order == ORC_ORDER_REFERENCE) {
    double t = 0.0;
    for (int64_t i = 0; i < N; ++i) t = t + w[i];
    return t;
  }
  int sh = weight_shift(N, w);
  u128 e = 0;
  for (int64_t i = 0; i < N; ++i) e += fixq(w[i], 96 - sh);
  return dbl128(e, 96 - sh);
}

// ll increment and ESS of stepFilter (model/ParticleFilter.scala:127-128, :431-434, :522-524)
void orc_ll_ess(int64_t N, const double* w1, double mx, int order, double* ll_incr, int32_t* ess) {
  double total = orc_total(N, w1, order);
  *ll_incr = mx + std::log(total / (double)N);
  double s2;
  if (order == ORC_ORDER_REFERENCE) {
    s2 = 0.0;
    for (int64_t i = 0; i < N; ++i) {
      double wn = w1[i] / total;
      s2 = s2 + wn * wn;
    }
  } else {
    // sum of squares exact (weights pre-scaled to <= 1 by a power of two), then one division by
    // total^2: sum (w/total)^2 = (sum w^2) / total^2
    int sh = weight_shift(N, w1);
    double sc = std::ldexp(1.0, -sh);
    const int q2 = (order == ORC_ORDER_DEVICE_F32) ? 48 : 96;  // F32 filters: 64-bit partial sums of w^2
    u128 e = 0;
    for (int64_t i = 0; i < N; ++i) {
      volatile double ws = w1[i] * sc;
      volatile double sq = ws * ws;  // exact for fp32-valued weights
      e += fixq(sq, q2);
    }
    volatile double ts = total * sc;
    volatile double tt = ts * ts;
    s2 = dbl128(e, q2) / tt;
  }
  double inv = std::floor(1 / s2);
  *ess = (inv != inv) ? 0 : (inv >= 2147483647.0 ? 2147483647 : (int32_t)inv);  // Scala .toInt saturates
}

// a9  Resampling (model/Resampling.scala:36-96).
//   systematic: k_i = (u + i)/n, one u (:63-72); stratified: k_i = (i + u_i)/n (:78-86);
//   both: treeEcdf (:52-58) = normalise, inclusive cumulative sum, TreeMap keyed by the sum --
//   a duplicated key keeps the LAST particle inserted -- and findAllInTreeMap (:36-46) = first
//   key >= k.  Where the reference would throw (k above the last key, m.head on an empty map)
//   the last particle is returned and *n_clamped counts it.
//   multinomial (:92-96): N independent Multinomial(w).draw; Breeze 1.0's first draw walks
//   prob = u*sum; for i: prob -= w_i; if (prob <= 0) return i.
//   ancestors are int32 indices into the input vector.
//   order | ORC_TIE_FIRST (8): the textbook inverse CDF instead (first index whose cumulative weight reaches k, no
//   duplicate-key rule) -- NOT the reference; the checker of the library's CSSM_TIE_FIRST option.
int orc_resample(int kind, int order, int64_t N, const double* w, const double* u, int32_t* anc,
                 int64_t* n_clamped) {
  const bool tie_first = (order & 8) != 0;
  order &= 7;
  int64_t clamped = 0;
  if (kind == CSSM_RESAMPLE_MULTINOMIAL) {
    if (order == ORC_ORDER_REFERENCE) {
      double sum = 0.0;
      for (int64_t i = 0; i < N; ++i) sum += w[i];
      for (int64_t o = 0; o < N; ++o) {
        double prob = u[o] * sum;
        int64_t j = -1;
        for (int64_t i = 0; i < N; ++i) {
          prob -= w[i];
          if (prob <= 0) { j = i; break; }
        }
        if (j < 0) { j = 0; ++clamped; }  // params.activeKeysIterator.next()
        anc[o] = (int32_t)j;
      }
    } else {
      int sh = weight_shift(N, w);
      std::vector<double> C(N);
      u128 e = 0;
      for (int64_t i = 0; i < N; ++i) { e += fixq(w[i], 96 - sh); C[i] = dbl128(e, 96 - sh); }
      double sum = C[N - 1];
      for (int64_t o = 0; o < N; ++o) {
        double target = u[o] * sum;
        int64_t j = std::lower_bound(C.begin(), C.end(), target) - C.begin();
        if (j >= N) { j = 0; ++clamped; }
        anc[o] = (int32_t)j;
      }
    }
    if (n_clamped) *n_clamped = clamped;
    return 0;
  }
  if (order != ORC_ORDER_REFERENCE) {
    // P_j = dbl128(exact cumulative sum), keys compared in the un-normalised domain
    int sh = weight_shift(N, w);
    std::vector<double> P(N);
    u128 e = 0;
    for (int64_t i = 0; i < N; ++i) { e += fixq(w[i], 96 - sh); P[i] = dbl128(e, 96 - sh); }
    const double total = P[N - 1], n = (double)N;
    int64_t j = 0;
    for (int64_t i = 0; i < N; ++i) {
      volatile double k = (kind == CSSM_RESAMPLE_SYSTEMATIC) ? (u[0] + (double)i) / n : ((double)i + u[i]) / n;
      volatile double target = k * total;
      while (j < N && P[j] < target) ++j;  // first key >= k
      if (j >= N) { anc[i] = (int32_t)(N - 1); ++clamped; continue; }
      // duplicate key: last insert wins.  In the reference a key repeats exactly when adding the
      // next normalised weight does not change the running sum, fl(C_j + wn_{j+1}) == C_j; the
      // device applies that same test to its own C_j = fl(P_j / total), so a run of vanishing
      // weights is skipped as a whole in both orders.
      int64_t jj = j;
      while (!tie_first && jj + 1 < N) {
        volatile double c = P[jj] / total;
        volatile double wn = w[jj + 1] / total;
        volatile double nx = c + wn;
        if (nx != c) break;
        ++jj;
      }
      anc[i] = (int32_t)jj;
    }
    if (n_clamped) *n_clamped = clamped;
    return 0;
  }
  // reference order: cumulative sums of the normalised weights
  std::vector<double> C(N), wn(N);
  double total = orc_total(N, w, order);
  for (int64_t i = 0; i < N; ++i) wn[i] = w[i] / total;
  {
    double c = 0.0;
    for (int64_t i = 0; i < N; ++i) { c = c + wn[i]; C[i] = c; }
  }
  // TreeMap semantics by a merge over the two sorted sequences
  int64_t j = 0;
  double n = (double)N;
  for (int64_t i = 0; i < N; ++i) {
    double k = (kind == CSSM_RESAMPLE_SYSTEMATIC) ? (u[0] + (double)i) / n : ((double)i + u[i]) / n;
    while (j < N && C[j] < k) ++j;  // first key >= k
    if (j >= N) { anc[i] = (int32_t)(N - 1); ++clamped; continue; }
    // duplicate key: last insert wins (model/Resampling.scala:55-57)
    int64_t jj = j;
    while (!tie_first && jj + 1 < N && C[jj + 1] == C[jj]) ++jj;
    anc[i] = (int32_t)jj;
  }
  if (n_clamped) *n_clamped = clamped;
  return 0;
}

// literal TreeMap version of the same thing (std::map = red-black tree, like
// scala.collection.immutable.TreeMap) -- used to validate the merge above and as the
// "reference-faithful" CPU-baseline variant
int orc_resample_treemap(int kind, int64_t N, const double* w, const double* u, int32_t* anc) {
  double total = 0.0;
  for (int64_t i = 0; i < N; ++i) total = total + w[i];
  std::map<double, int32_t> ecdf;
  double c = 0.0;
  for (int64_t i = 0; i < N; ++i) { c = c + w[i] / total; ecdf[c] = (int32_t)i; }
  double n = (double)N;
  auto it = ecdf.begin();
  for (int64_t i = 0; i < N; ++i) {
    double k = (kind == CSSM_RESAMPLE_SYSTEMATIC) ? (u[0] + (double)i) / n : ((double)i + u[i]) / n;
    it = ecdf.lower_bound(k);
    anc[i] = (it == ecdf.end()) ? (int32_t)(N - 1) : it->second;
  }
  return 0;
}

void orc_gather(int64_t N, int d, const double* x, const int32_t* anc, double* out) {
  for (int k = 0; k < d; ++k)
    for (int64_t i = 0; i < N; ++i) out[k * N + i] = x[k * N + anc[i]];
}

// a3  one stepFilter (model/ParticleFilter.scala:116-132) with injected noise z[d][N] and
//     resampling uniforms u.  Outputs may be NULL except x_out.  ll/ess are in-out.
int orc_step_filter(const cssm_model_desc_t* m, int64_t N, int resample_kind, int order,
                    double t_prev, double t, int has_obs, double y, const double* x_in,
                    const double* z, const double* u, double* x_prop, double* logw_out,
                    double* w1_out, int32_t* anc_out, double* x_out, double* ll, int32_t* ess) {
  int d = orc_dim(m);
  double dt = t - t_prev;
  std::vector<double> x1((size_t)d * N);
  orc_propagate(m, N, dt, x_in, z, x1.data());
  if (x_prop) std::memcpy(x_prop, x1.data(), sizeof(double) * d * N);
  if (!has_obs) {  // :121
    std::memcpy(x_out, x1.data(), sizeof(double) * d * N);
    return 0;
  }
  std::vector<double> g(N), w(N), w1(N);
  std::vector<int32_t> anc(N);
  orc_f(m, N, x1.data(), t, g.data());
  orc_loglik(m, N, g.data(), y, w.data());
  double mx = orc_max(N, w.data());
  orc_w1(N, w.data(), mx, order, w1.data());
  orc_resample(resample_kind, order, N, w1.data(), u, anc.data(), nullptr);
  double incr;
  orc_ll_ess(N, w1.data(), mx, order, &incr, ess);
  *ll = *ll + incr;  // s.ll + max + log(mean(w1)), :127
  orc_gather(N, d, x1.data(), anc.data(), x_out);
  if (logw_out) std::memcpy(logw_out, w.data(), sizeof(double) * N);
  if (w1_out) std::memcpy(w1_out, w1.data(), sizeof(double) * N);
  if (anc_out) std::memcpy(anc_out, anc.data(), sizeof(int32_t) * N);
  return 0;
}

// number of sub-steps of FilterLgcp.calcWeight (model/ParticleFilter.scala:190)
int64_t orc_lgcp_nsub(double dt, int precision) {
  if (dt == 0) return 0;
  return (int64_t)(int)std::ceil(dt / std::pow(10, -precision));
}

// a10 FilterLgcp.stepFilter (model/ParticleFilter.scala:184-226): n sub-steps of the SDE's
//     stepFunction(delta), delta = 10^-precision; the stream starts at time y.t (sic, :215,:194)
//     and its first element is already one step past the start (MarkovChain.draw,
//     model/MarkovChain.scala:7-12); hazard = sum exp(f(x_i, t_i)) * delta over the n post-step
//     states; gamma = f(x_n, y.t); log-weight = gamma - hazard; always resamples.
//     z: [n_sub][d][N]
int orc_step_lgcp(const cssm_model_desc_t* m, int64_t N, int resample_kind, int order, double t_prev,
                  double t, const double* x_in, const double* z, const double* u, double* x_prop,
                  double* logw_out, double* w1_out, int32_t* anc_out, double* x_out, double* ll,
                  int32_t* ess) {
  int d = orc_dim(m);
  double dt = t - t_prev;
  double delta = std::pow(10, -m->lgcp_precision);
  int64_t n = orc_lgcp_nsub(dt, m->lgcp_precision);
  std::vector<double> xa(x_in, x_in + (size_t)d * N), xb((size_t)d * N), w(N), w1(N), g(N);
  std::vector<int32_t> anc(N);
  if (dt == 0) {  // :212-213  weight = f - f
    orc_f(m, N, xa.data(), t, g.data());
    for (int64_t i = 0; i < N; ++i) w[i] = g[i] - g[i];
  } else {
    std::vector<double> hz(N, 0.0);
    double time = t;  // simInitStream(t, x, delta): t0 of the stream is y.t
    for (int64_t s = 0; s < n; ++s) {
      orc_propagate(m, N, delta, xa.data(), z + (size_t)s * d * N, xb.data());
      xa.swap(xb);
      time = time + delta;
      orc_f(m, N, xa.data(), time, g.data());
      for (int64_t i = 0; i < N; ++i) hz[i] = hz[i] + std::exp(g[i]) * delta;
    }
    orc_f(m, N, xa.data(), t, g.data());
    for (int64_t i = 0; i < N; ++i) w[i] = g[i] - hz[i];
  }
  if (x_prop) std::memcpy(x_prop, xa.data(), sizeof(double) * d * N);
  double mx = orc_max(N, w.data());
  orc_w1(N, w.data(), mx, order, w1.data());
  double incr;
  orc_ll_ess(N, w1.data(), mx, order, &incr, ess);
  *ll = *ll + incr;
  orc_resample(resample_kind, order, N, w1.data(), u, anc.data(), nullptr);
  orc_gather(N, d, xa.data(), anc.data(), x_out);
  if (logw_out) std::memcpy(logw_out, w.data(), sizeof(double) * N);
  if (w1_out) std::memcpy(w1_out, w1.data(), sizeof(double) * N);
  if (anc_out) std::memcpy(anc_out, anc.data(), sizeof(int32_t) * N);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// self-driven runs (own RNG): Monte-Carlo agreement tests and the CPU baseline timing
// ---------------------------------------------------------------------------------------------
struct Rng {  // xoshiro256++ with a polar-method normal
  uint64_t s[4];
  bool have;
  double spare;
  explicit Rng(uint64_t seed) : have(false), spare(0) {
    uint64_t z = seed;
    for (int i = 0; i < 4; ++i) {  // splitmix64
      z += 0x9E3779B97F4A7C15ull;
      uint64_t v = z;
      v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
      v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
      s[i] = v ^ (v >> 31);
    }
  }
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  inline uint64_t next() {
    uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  inline double normal() {
    if (have) { have = false; return spare; }
    double a, b, r;
    do { a = 2 * uniform() - 1; b = 2 * uniform() - 1; r = a * a + b * b; } while (r >= 1 || r == 0);
    double f = std::sqrt(-2 * std::log(r) / r);
    spare = b * f; have = true;
    return a * f;
  }
};

// llFilter (model/ParticleFilter.scala:137-140) with the oracle's own generator.
//   variant 0 "faithful": particle-major (AoS) storage, everything recomputed per particle as the
//             reference does, std::map ECDF + lower_bound per output (TreeMap, :36-58);
//   variant 1 "flat": SoA, merge-based ancestor search -- a fair lower bound for a tuned CPU code.
//   Works for every obs kind incl. LGCP (through orc_step_lgcp).  ess_out / ll_steps_out: T or NULL.
double orc_filter_ll(const cssm_model_desc_t* m, int64_t N, int resample_kind, int64_t T,
                     const double* t, const double* y, const uint8_t* has_obs, uint64_t seed,
                     int variant, double* ll_steps_out, int32_t* ess_out) {
  int d = orc_dim(m);
  Rng rng(seed);
  std::vector<double> x((size_t)d * N), x2((size_t)d * N), z((size_t)d * N), u(N), g(N), w(N), w1(N);
  std::vector<int32_t> anc(N);
  for (auto& v : z) v = rng.normal();
  orc_init_state(m, N, z.data(), x.data());
  double t0 = t[0];
  for (int64_t s = 1; s < T; ++s) t0 = std::min(t0, t[s]);
  double ll = 0.0, tp = t0;
  int32_t ess = (int32_t)N;
  for (int64_t s = 0; s < T; ++s) {
    int nu = (resample_kind == CSSM_RESAMPLE_SYSTEMATIC) ? 1 : (int)N;
    if (m->obs_kind == CSSM_OBS_LGCP) {
      int64_t n = orc_lgcp_nsub(t[s] - tp, m->lgcp_precision);
      std::vector<double> zz((size_t)n * d * N);
      for (auto& v : zz) v = rng.normal();
      for (int i = 0; i < nu; ++i) u[i] = rng.uniform();
      orc_step_lgcp(m, N, resample_kind, ORC_ORDER_REFERENCE, tp, t[s], x.data(), zz.data(), u.data(),
                    nullptr, nullptr, nullptr, nullptr, x2.data(), &ll, &ess);
      x.swap(x2);
    } else if (variant == 1) {
      for (auto& v : z) v = rng.normal();
      for (int i = 0; i < nu; ++i) u[i] = rng.uniform();
      orc_step_filter(m, N, resample_kind, ORC_ORDER_REFERENCE, tp, t[s], has_obs ? has_obs[s] : 1, y[s],
                      x.data(), z.data(), u.data(), nullptr, nullptr, nullptr, nullptr, x2.data(), &ll, &ess);
      x.swap(x2);
    } else {
      // faithful cost model: particle by particle, TreeMap ECDF
      double dt = t[s] - tp;
      std::vector<std::vector<double>> parts(N, std::vector<double>(d)), parts2;
      for (int64_t i = 0; i < N; ++i)
        for (int k = 0; k < d; ++k) parts[i][k] = x[k * N + i];
      for (int64_t i = 0; i < N; ++i) {
        int k = 0;
        for (int l = 0; l < m->n_leaves; ++l)
          for (int c = 0; c < m->leaves[l].dim; ++c, ++k)
            parts[i][k] = (m->step_mode == CSSM_STEP_EULER)
                              ? step_euler_1(m->leaves[l], c, dt, parts[i][k], rng.normal())
                              : step_exact_1(m->leaves[l], c, dt, parts[i][k], rng.normal());
      }
      for (int64_t i = 0; i < N; ++i)
        for (int k = 0; k < d; ++k) x[k * N + i] = parts[i][k];
      if (!has_obs || has_obs[s]) {
        for (int64_t i = 0; i < N; ++i) w[i] = orc_loglik_1(m, f_one(m, N, x.data(), i, t[s]), y[s]);
        double mx = orc_max(N, w.data());
        for (int64_t i = 0; i < N; ++i) w1[i] = std::exp(w[i] - mx);
        for (int i = 0; i < nu; ++i) u[i] = rng.uniform();
        if (resample_kind == CSSM_RESAMPLE_MULTINOMIAL)
          orc_resample(resample_kind, ORC_ORDER_REFERENCE, N, w1.data(), u.data(), anc.data(), nullptr);
        else
          orc_resample_treemap(resample_kind, N, w1.data(), u.data(), anc.data());
        double incr;
        orc_ll_ess(N, w1.data(), mx, ORC_ORDER_REFERENCE, &incr, &ess);
        ll = ll + incr;
        orc_gather(N, d, x.data(), anc.data(), x2.data());
        x.swap(x2);
      }
    }
    tp = t[s];
    if (ll_steps_out) ll_steps_out[s] = ll;
    if (ess_out) ess_out[s] = ess;
  }
  return ll;
}

// R independent filters on `threads` host threads (the reference's only concurrency is across
// independent filters: Streaming.pilotRun mapAsyncUnordered(4), model/Streaming.scala:39; PMMH
// chains mapAsync(2), examples/DetermineParameters.scala:69).  ll_out[R].
void orc_filter_ll_many(const cssm_model_desc_t* m, int64_t N, int resample_kind, int64_t T,
                        const double* t, const double* y, const uint8_t* has_obs, uint64_t seed,
                        int variant, int R, int threads, double* ll_out) {
  std::vector<std::thread> pool;
  for (int th = 0; th < threads; ++th)
    pool.emplace_back([=]() {
      for (int r = th; r < R; r += threads)
        ll_out[r] = orc_filter_ll(m, N, resample_kind, T, t, y, has_obs, seed + 1000003ull * r, variant, nullptr, nullptr);
    });
  for (auto& p : pool) p.join();
}

// SimulateData.simStep (model/Data.scala:186-193) on a regular grid from t = 0
// (simMarkov, :81-91): used ONLY to make synthetic observations.  Observation samplers follow
// model/Model.scala:169-178 (NegBin as Gamma-Poisson), :267 (Poisson), :212,245 (Normal),
// :316 (Bernoulli).  Simple inversion / Knuth samplers; the draws need not match Breeze's.
static double draw_poisson(Rng& r, double lam) {
  if (lam < 30) {
    double L = std::exp(-lam), p = 1;
    int k = 0;
    do { ++k; p *= r.uniform(); } while (p > L);
    return k - 1;
  }
  double v = lam + std::sqrt(lam) * r.normal();  // normal approximation for large means
  return v < 0 ? 0 : std::floor(v + 0.5);
}
static double draw_gamma(Rng& r, double shape, double scale) {  // Marsaglia-Tsang
  if (shape < 1) return draw_gamma(r, shape + 1, scale) * std::pow(r.uniform(), 1 / shape);
  double dd = shape - 1.0 / 3, c = 1 / std::sqrt(9 * dd);
  for (;;) {
    double xx = r.normal(), v = 1 + c * xx;
    if (v <= 0) continue;
    v = v * v * v;
    double uu = r.uniform();
    if (std::log(uu) < 0.5 * xx * xx + dd - dd * v + dd * std::log(v)) return dd * v * scale;
  }
}
void orc_simulate(const cssm_model_desc_t* m, int64_t T, double dt, uint64_t seed, double* t_out,
                  double* y_out, double* x_out /* [T][d] or NULL */) {
  int d = orc_dim(m);
  Rng rng(seed);
  std::vector<double> x(d), z(d), x2(d);
  for (auto& v : z) v = rng.normal();
  orc_init_state(m, 1, z.data(), x.data());
  double t = 0.0;
  for (int64_t s = 0; s < T; ++s) {
    if (s > 0) {
      for (auto& v : z) v = rng.normal();
      orc_propagate(m, 1, dt, x.data(), z.data(), x2.data());
      x.swap(x2);
      t = t + dt;
    }
    double g = f_one(m, 1, x.data(), 0, t), yv = 0;
    switch (m->obs_kind) {
      case CSSM_OBS_POISSON: yv = draw_poisson(rng, std::exp(g)); break;
      case CSSM_OBS_NEGBIN: {
        double size = std::exp(m->scale), mu = std::exp(g), prob = mu / (size + mu);
        yv = draw_poisson(rng, draw_gamma(rng, size, prob / (1 - prob)));
        break;
      }
      case CSSM_OBS_NORMAL: yv = g + std::exp(m->scale) * rng.normal(); break;
      case CSSM_OBS_BERNOULLI: {
        double p = (g > 6) ? 1.0 : (g < -6) ? 0.0 : 1.0 / (1 + std::exp(-g));
        yv = rng.uniform() < p ? 1.0 : 0.0;
        break;
      }
      case CSSM_OBS_STUDENT_T: {  // StudentsT(df) * v + x (:145-150): normal / sqrt(chi2_df / df)
        double chi2 = 0.0;
        for (int i = 0; i < m->obs_df; ++i) { double n = rng.normal(); chi2 += n * n; }
        yv = rng.normal() / std::sqrt(chi2 / m->obs_df) * std::exp(m->scale) + g;
        break;
      }
      case CSSM_OBS_ZIP: {  // :282-292
        double p = std::exp(m->scale) / (1 + std::exp(m->scale));
        double nz = draw_poisson(rng, std::exp(g));
        yv = rng.uniform() < p ? 0.0 : nz;
        break;
      }
      case CSSM_OBS_BETA: {  // Beta(exp(-gamma), beta) (:340-343) as a ratio of gammas, kept inside (0, 1)
        double ga = draw_gamma(rng, std::exp(-g), 1.0), gb = draw_gamma(rng, m->has_scale ? m->scale : 1.0, 1.0);
        yv = std::min(std::max(ga / (ga + gb), 1e-12), 1.0 - 1e-12);
        break;
      }
      default: yv = 1.0; break;
    }
    t_out[s] = t;
    y_out[s] = yv;
    if (x_out) std::memcpy(x_out + s * d, x.data(), sizeof(double) * d);
  }
}

}  // extern "C"
