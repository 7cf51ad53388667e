// cssm_common.cuh -- device-side building blocks shared by all kernels of libcssm_gpu.so
// (sm_100a only).  No reference code is involved here: 128-bit fixed point, the deterministic
// exp, Philox4x32-10 and warp/block helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cssm {

// ---------------------------------------------------------------------------------------------
// 128-bit unsigned fixed point (value * 2^-q).  Integer addition is associative, so any scan /
// reduction order -- any tile size, grid size or GPU count -- gives the same bits.
// ---------------------------------------------------------------------------------------------
struct u128 {
  unsigned long long lo, hi;
};
__host__ __device__ __forceinline__ u128 make_u128(unsigned long long lo, unsigned long long hi) {
  u128 r;
  r.lo = lo;
  r.hi = hi;
  return r;
}
__host__ __device__ __forceinline__ u128 add128(u128 a, u128 b) {
  u128 r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u64 %0, %2, %4;\n\taddc.u64 %1, %3, %5;" : "=l"(r.lo), "=l"(r.hi) : "l"(a.lo), "l"(a.hi), "l"(b.lo), "l"(b.hi));
#else
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
#endif
  return r;
}
__host__ __device__ __forceinline__ bool is_zero128(u128 a) { return (a.lo | a.hi) == 0ull; }

// floor(x * 2^q), x >= 0 finite with x * 2^q < 2^100 (callers keep x <= 1, q <= 96);
// zero / negative / NaN / subnormal -> 0
__host__ __device__ __forceinline__ u128 fixq(double x, int q) {
  if (!(x > 0.0)) return make_u128(0, 0);
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
  unsigned long long b;
  memcpy(&b, &x, 8);
#endif
  int ef = (int)((b >> 52) & 0x7ff);
  if (ef == 0) return make_u128(0, 0);
  unsigned long long mant = (b & 0xFFFFFFFFFFFFFull) | (1ull << 52);
  int sh = ef - 1075 + q;
  if (sh >= 0) {
    if (sh == 0) return make_u128(mant, 0);
    if (sh < 64) return make_u128(mant << sh, mant >> (64 - sh));
    return make_u128(0, mant << (sh - 64));
  }
  if (sh <= -53) return make_u128(0, 0);
  return make_u128(mant >> (-sh), 0);
}

// round-to-nearest-even of e * 2^-q  (q in [32, 128))
__host__ __device__ __forceinline__ double unfixq(u128 e, int q) {
  if (is_zero128(e)) return 0.0;
  int p;
#ifdef __CUDA_ARCH__
  p = e.hi ? 127 - __clzll((long long)e.hi) : 63 - __clzll((long long)e.lo);
#else
  p = e.hi ? 127 - __builtin_clzll(e.hi) : 63 - __builtin_clzll(e.lo);
#endif
  unsigned long long mant;
  int s = 0;
  if (p <= 52) {
    mant = e.lo;
  } else {
    s = p - 52;  // 1..75
    unsigned long long rem_hi, rem_lo, half_hi, half_lo;
    if (s < 64) {
      mant = (e.lo >> s) | (e.hi << (64 - s));  // s >= 1 so the shift is < 64; e.hi < 2^(p-63)
      rem_hi = 0;
      rem_lo = e.lo & ((1ull << s) - 1ull);
      half_hi = 0;
      half_lo = 1ull << (s - 1);
    } else {
      int s2 = s - 64;  // 0..11
      mant = e.hi >> s2;
      rem_hi = s2 ? (e.hi & ((1ull << s2) - 1ull)) : 0ull;
      rem_lo = e.lo;
      half_hi = s2 ? (1ull << (s2 - 1)) : 0ull;
      half_lo = s2 ? 0ull : (1ull << 63);
    }
    bool gt = (rem_hi > half_hi) || (rem_hi == half_hi && rem_lo > half_lo);
    bool eq = (rem_hi == half_hi) && (rem_lo == half_lo);
    if (gt || (eq && (mant & 1ull))) mant += 1ull;
  }
  // (double)mant is exact (mant <= 2^53); the power of two is normal: s - q in [-127, 43]
  unsigned long long sb = (unsigned long long)(s - q + 1023) << 52;
#ifdef __CUDA_ARCH__
  return __dmul_rn((double)mant, __longlong_as_double((long long)sb));
#else
  double sc;
  memcpy(&sc, &sb, 8);
  return (double)mant * sc;
#endif
}

#ifdef __CUDACC__
// floor(w * 2^q) for 0 <= w with w * 2^q < 2^128, q >= 64: the same integer as fixq, by two exact
// truncating conversions (scaling by a power of two and taking the fractional part are exact).
// NaN and negative inputs give 0 (cvt.rzi.u64.f64 saturates).
__device__ __forceinline__ u128 fix_fast(double w, int q) {
  const double a = __dmul_rn(w, __longlong_as_double((long long)(q - 64 + 1023) << 52));
  const unsigned long long hi = __double2ull_rz(a);
  const double r = __dsub_rn(a, __ull2double_rn(hi));  // exact: hi <= a < 2^53 * ulp
  const unsigned long long lo = __double2ull_rz(__dmul_rn(r, 18446744073709551616.0));
  return make_u128(lo, hi);
}
// the device's conversion of an exact sum to fp64: (double)hi * 2^64 + (double)lo, scaled by 2^-q.
// Two correctly rounded conversions and one add: bit-identical to oracle/cssm_oracle.cpp dbl128,
// monotone in e, within 1.5 ulp of e * 2^-q.
__device__ __forceinline__ double dbl128(u128 e, int q) {
  const double c = __dadd_rn(__dmul_rn(__ull2double_rn(e.hi), 18446744073709551616.0), __ull2double_rn(e.lo));
  return __dmul_rn(c, __longlong_as_double((long long)(1023 - q) << 52));
}

// ---------------------------------------------------------------------------------------------
// Deterministic exp: the same fma/mul/add sequence as oracle/cssm_oracle.cpp orc_exp_det /
// orc_expf_det, so the weights w1 = exp(logw - max) have identical bits on the device and in the
// checker.  Coefficients live in constant memory so that every DFMA/FFMA takes its constant as a
// direct operand (the compiler otherwise re-materialises each 64-bit literal per use).
// ---------------------------------------------------------------------------------------------
__constant__ double c_expd[17] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0,       1.0 / 720.0,       1.0 / 120.0,       1.0 / 24.0,      1.0 / 6.0,      0.5,
    1.4426950408889634074 /* log2 e */, -6.93147180369123816490e-01 /* -ln2 hi */, -1.90821492927058770002e-10 /* -ln2 lo */,
    6755399441055744.0 /* 1.5 * 2^52 */, 18446744073709551616.0 /* 2^64 */};
__constant__ float c_expf[11] = {1.0f / 5040.0f, 1.0f / 720.0f, 1.0f / 120.0f, 1.0f / 24.0f, 1.0f / 6.0f, 0.5f,
                                 1.442695040888963f /* log2 e */, -0.693145751953125f /* -ln2 hi (0x3f317200) */,
                                 -1.428606765330187e-06f /* -ln2 lo */, 12582912.0f /* 1.5 * 2^23 */, 0.0f};

// fp64 weights (F64 filters, cssm_resample).  |error| < 1 ulp.
__device__ __forceinline__ double exp_det(double x) {
  if (!(x >= -745.5)) return (x != x) ? x : 0.0;
  // k = rint(x * log2 e) by the magic-number add (round to nearest even, like nearbyint)
  const double t = __dadd_rn(__dmul_rn(x, c_expd[12]), c_expd[15]);
  const double kf = __dsub_rn(t, c_expd[15]);
  const int k = __double2loint(t);  // low word of 1.5*2^52 + k is k in two's complement
  double r = __fma_rn(kf, c_expd[13], x);
  r = __fma_rn(kf, c_expd[14], r);
  double p = c_expd[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = __fma_rn(p, r, c_expd[i]);
  p = __fma_rn(p, r, 1.0);
  p = __fma_rn(p, r, 1.0);
  if (k > -1000)  // p in (0.7, 1.42): scaling by 2^k is an exponent add while the result is normal
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  // near the subnormal range: scale in two exact steps so that the result rounds once
  const int k1 = k / 2, k2 = k - k1;
  const double s1 = __longlong_as_double((long long)(k1 + 1023) << 52);
  const double s2 = __longlong_as_double((long long)(k2 + 1023) << 52);
  return __dmul_rn(__dmul_rn(p, s1), s2);
}

// fp32 weights (F32 filters): everything on the full-rate fp32 / integer pipes.  Arguments below
// -86 give exactly 0 (the result would be subnormal in fp32); |error| < 1 ulp(fp32) above.
__device__ __forceinline__ float expf_det(float x) {
  const float t = __fadd_rn(__fmul_rn(x, c_expf[6]), c_expf[9]);
  const float kf = __fsub_rn(t, c_expf[9]);
  const int k = (int)(__float_as_uint(t) << 10) >> 10;  // low 22 bits of 1.5*2^23 + k, sign-extended
  float r = __fmaf_rn(kf, c_expf[7], x);
  r = __fmaf_rn(kf, c_expf[8], r);
  float p = c_expf[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) p = __fmaf_rn(p, r, c_expf[i]);
  p = __fmaf_rn(p, r, 1.0f);
  p = __fmaf_rn(p, r, 1.0f);
  const float v = __uint_as_float(__float_as_uint(p) + ((unsigned)k << 23));  // k >= -124: the result is normal
  return (x >= -86.0f) ? v : ((x != x) ? x : 0.0f);  // selects, no branch: the loads of a tile stay batched
}

// floor(w * 2^96) of an fp32 weight 0 <= w <= 1: two exact truncating conversions, as fix_fast.
// hi = floor(w * 2^32); when w * 2^32 >= 2^24 it is an integer and the remainder is 0, otherwise
// the remainder w * 2^32 - hi is exact in fp32.  NaN -> 0.
__device__ __forceinline__ u128 fix_f32(float w) {
  const float a = __fmul_rn(w, 4294967296.0f);
  const unsigned long long hi = __float2ull_rz(a);
  const float r = __fsub_rn(a, __ull2float_rn(hi));
  const unsigned long long lo = __float2ull_rz(__fmul_rn(r, 18446744073709551616.0f));
  return make_u128(lo, hi);
}
// PTX shifts clamp the amount at the register width (an amount >= 64, or a negative one read as
// unsigned, gives 0), which is exactly what a 128-bit shift assembled from 64-bit pieces needs
__device__ __forceinline__ unsigned long long shl64c(unsigned long long x, int n) {
  unsigned long long r;
  asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n));
  return r;
}
__device__ __forceinline__ unsigned long long shr64c(unsigned long long x, int n) {
  unsigned long long r;
  asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n));
  return r;
}
// floor(w*w * 2^48) of an fp32 weight 0 <= w <= 1 (the ESS accumulator of F32 filters; fits 64 bits).
// (double)w * 2^24 and its square are exact in fp64 (48-bit product), the truncating conversion is
// the floor: the same integer as m^2 * 2^(2(e-150)+48) assembled from the mantissa, but on the
// conversion / fp64 pipes that the weight pass leaves idle instead of a dozen integer instructions.
// NaN, zero and subnormal weights give 0.
__device__ __forceinline__ unsigned long long fix_sq48_f32(float w) {
  const double a = __dmul_rn((double)w, 16777216.0);
  return __double2ull_rz(__dmul_rn(a, a));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32 (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3"), counter-based, all state in registers.
// Rounds: 7 is the smallest count for which the paper reports Philox4x32 to pass the complete BigCrush battery
// ("Crush-resistant"); 10 is the Random123 / cuRAND default with extra margin.  The generator is 35 % of K1's
// instructions, and K1 is bound by instruction issue: 7 rounds take it from 0.206 to 0.191 ms at 2^24 particles (78 % ->
// 87 % of the HBM peak, profiles/r02_summary.md).  -DCSSM_PHILOX_ROUNDS=10 builds the conservative variant.
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
};
#ifndef CSSM_PHILOX_ROUNDS
#define CSSM_PHILOX_ROUNDS 7
#endif
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < CSSM_PHILOX_ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// what a counter block is used for (upper byte of counter word 3)
enum : uint32_t { RNG_INIT = 1u << 24, RNG_STEP = 2u << 24, RNG_RESAMPLE = 3u << 24, RNG_SAMPLE_ONE = 4u << 24 };

__device__ __forceinline__ double u64_to_unit_double(uint32_t a, uint32_t b) {
  unsigned long long v = ((unsigned long long)a << 32) | b;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
}

// two uint32 -> two N(0,1) floats (Box-Muller entirely on the SFU: lg2, sqrt, sin, cos approx)
__device__ __forceinline__ void box_muller_f(uint32_t a, uint32_t b, float& z0, float& z1) {
  float u1 = fmaf(__uint2float_rz(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0,1)
  float ang = __uint2float_rz(b) * 1.4629180792671596e-09f;                                 // 2 pi * [0,1)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(u1)));
  z0 = r * __cosf(ang);
  z1 = r * __sinf(ang);
}
// four uint32 -> two N(0,1) doubles
__device__ __forceinline__ void box_muller_d(uint4 v, double& z0, double& z1) {
  double u1 = u64_to_unit_double(v.x, v.y);
  double u2 = u64_to_unit_double(v.z, v.w);
  u1 = 1.0 - u1;  // (0, 1]
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// ---------------------------------------------------------------------------------------------
// ordered-integer view of a double so that max can be taken with one integer atomic
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long ord_key(double v) {
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ord_unkey(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

__device__ __forceinline__ u128 shfl_xor128(u128 v, int m) {
  u128 r;
  r.lo = __shfl_xor_sync(0xffffffffu, v.lo, m);
  r.hi = __shfl_xor_sync(0xffffffffu, v.hi, m);
  return r;
}
__device__ __forceinline__ u128 shfl_up128(u128 v, int d) {
  u128 r;
  r.lo = __shfl_up_sync(0xffffffffu, v.lo, d);
  r.hi = __shfl_up_sync(0xffffffffu, v.hi, d);
  return r;
}
__device__ __forceinline__ u128 warp_sum128(u128 v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = add128(v, shfl_xor128(v, m));
  return v;
}
// exact 128-bit atomic add built from two 64-bit atomics; commutative, so the final value does
// not depend on the order in which blocks arrive
__device__ __forceinline__ void atomic_add128(u128* dst, u128 v) {
  unsigned long long old = atomicAdd(&dst->lo, v.lo);
  unsigned long long carry = (old + v.lo < old) ? 1ull : 0ull;
  if (v.hi | carry) atomicAdd(&dst->hi, v.hi + carry);
}

// ---------------------------------------------------------------------------------------------
// flags and small payloads exchanged between the ranks of a sharded filter.  The slots live in
// the DESTINATION rank's memory and are written by peers over NVLink (P2P mapped pointers).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// L2-coherent read of a value other blocks of this grid updated with atomics
__device__ __forceinline__ unsigned long long ld_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// tagged 64-bit words (tag << 32 | payload) handed from a producer block to a consumer block of the same grid through L2:
// an aligned 8-byte store is single-copy atomic, so a word whose tag matches carries its payload -- no fence, no flag
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_gpu_pair(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// programmatic dependent launch: wait for the preceding kernel of the stream / let the next start
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif  // __CUDACC__

}  // namespace cssm
