// CssmNative.scala -- @native declarations bound by jvm/cssm_jni.c.  NOT COMPILED IN THIS IMAGE (no JVM); the C side is
// compiled and exercised through a fake JNIEnv by tests/test_jni_shim.py, which also checks that every @native method
// below has its Java_com_github_jonnylaw_gpu_CssmNative_00024_<name> entry point and vice versa.
package com.github.jonnylaw.gpu

object CssmNative {
  System.loadLibrary("cssm_jni")   // which links libcssm_gpu.so
  @native def filterCreate(kinds: Array[Int], params: Array[Double], obsKind: Int, hasScale: Boolean, scale: Double,
    stepMode: Int, precision: Int, obsDf: Int, n: Long, resampleKind: Int, dtype: Int, device: Int, seed: Long,
    streamId: Long): Long
  @native def filterSetParams(h: Long, kinds: Array[Int], params: Array[Double], obsKind: Int, hasScale: Boolean,
    scale: Double, stepMode: Int, precision: Int, obsDf: Int): Unit
  @native def filterDestroy(h: Long): Unit
  @native def filterInit(h: Long, t0: Double): Unit
  @native def filterInitState(h: Long, t0: Double, x0: Array[Double]): Unit
  @native def filterStep(h: Long, t: Double, hasObs: Boolean, y: Double, essOut: Array[Int]): Double
  @native def filterLl(h: Long, t: Array[Double], y: Array[Double], hasObs: Array[Byte]): Double
  @native def filterRun(h: Long, t: Array[Double], y: Array[Double], hasObs: Array[Byte], statesOut: Array[Double]): Double
  /** the resident-series form PMMH uses: load once, then filterSetParams + filterLlResident per proposal */
  @native def filterLoadSeries(h: Long, t: Array[Double], y: Array[Double], hasObs: Array[Byte]): Unit
  @native def filterLlResident(h: Long): Double
  @native def filterSeriesLen(h: Long): Long
  /** 0 = the reference's TreeMap rule (default), 1 = first index (include/cssm.h, cssm_tie_rule) */
  @native def filterSetTieRule(h: Long, rule: Int): Unit
  @native def filterReseed(h: Long, seed: Long, streamId: Long): Unit
  @native def filterGetParticles(h: Long, out: Array[Double]): Unit
  @native def filterSeriesMode(h: Long, mode: Int): Unit
  /** out = mean[d] | lower[d] | upper[d] | gammaLower, gammaUpper (ParticleFilter.getIntervals on the device) */
  @native def filterIntervals(h: Long, t: Double, interval: Double, d: Int, out: Array[Double]): Unit
  @native def filterSampleOne(h: Long, out: Array[Double]): Unit
  /** out = mean[d] | lower[d] | upper[d] | eta mean, lower, upper | obs mean, lower, upper (getMeanForecast on the device) */
  @native def filterForecast(h: Long, t: Double, interval: Double, chain: Boolean, d: Int, out: Array[Double]): Unit
  /** out = x[d*N] | gamma[N] | eta[N] | obs[N] (getForecast) */
  @native def filterForecastCloud(h: Long, d: Int, n: Long, out: Array[Double]): Unit
  @native def filterPathsEnable(h: Long, maxSteps: Long): Unit
  @native def filterPathsLen(h: Long): Long
  /** out[idx.length][len + 1][d], oldest state first (FilterInterpolate) */
  @native def filterGetPaths(h: Long, idx: Array[Int], out: Array[Double]): Unit
  @native def resample(kind: Int, w: Array[Double], u: Array[Double], anc: Array[Int], device: Int): Unit
}
