/* cssm_jni.c -- JNI shim, one-to-one over include/cssm.h.
 *
 * On a machine with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       cssm_jni.c -L../composablestatespacemodels_b200/csrc -lcssm_gpu -o libcssm_jni.so
 * This image has no JDK: tests/test_jni_shim.py compiles this file with -Wall -Werror against the stand-in header
 * jvm/jni_stub/jni.h (same names and signatures as the JDK's), links it against libcssm_gpu.so and drives the
 * Java_... entry points through a fake JNIEnv (tests/jni_harness.c).
 * Java side: jvm/CssmNative.scala (`@native` methods of object CssmNative).
 * Every non-zero status is rethrown as RuntimeException(cssm_last_error()), which is how the
 * reference reports errors (thrown exceptions, model/Sde.scala:183, model/Model.scala:46).
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include "cssm.h"

static int check(JNIEnv* env, int status) {
  if (status != 0) {
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
    (*env)->ThrowNew(env, ex, cssm_last_error());
  }
  return status;
}

/* leaves arrive flattened: kinds[n_leaves*5] = (sde_kind, dim, f_kind, period, harmonics) per leaf and
 * params[5 * d] = m0 | c0 | phi | mu | sigma, each of total length d in leaf order */
static void build_desc(JNIEnv* env, jintArray kinds, jdoubleArray params, jint obs_kind, jboolean has_scale,
                       jdouble scale, jint step_mode, jint precision, jint obs_df, cssm_model_desc_t* desc,
                       cssm_leaf_t** leaves_out, jint** k_out, jdouble** p_out) {
  jsize nk = (*env)->GetArrayLength(env, kinds) / 5;
  jint* k = (*env)->GetIntArrayElements(env, kinds, NULL);
  jdouble* p = (*env)->GetDoubleArrayElements(env, params, NULL);
  jsize d = (*env)->GetArrayLength(env, params) / 5;
  cssm_leaf_t* leaves = (cssm_leaf_t*)calloc((size_t)nk, sizeof(cssm_leaf_t));
  int off = 0;
  for (jsize l = 0; l < nk; ++l) {
    leaves[l].sde_kind = k[5 * l]; leaves[l].dim = k[5 * l + 1]; leaves[l].f_kind = k[5 * l + 2];
    leaves[l].period = k[5 * l + 3]; leaves[l].harmonics = k[5 * l + 4];
    leaves[l].m0 = p + off; leaves[l].c0 = p + d + off; leaves[l].phi = p + 2 * d + off;
    leaves[l].mu = p + 3 * d + off; leaves[l].sigma = p + 4 * d + off;
    off += leaves[l].dim;
  }
  desc->n_leaves = (int32_t)nk; desc->leaves = leaves; desc->obs_kind = obs_kind; desc->has_scale = has_scale;
  desc->scale = scale; desc->step_mode = step_mode; desc->lgcp_precision = precision;
  desc->obs_df = obs_df; desc->reserved = 0; /* StudentsTModel.df, model/Model.scala:144 */
  *leaves_out = leaves; *k_out = k; *p_out = p;
}

JNIEXPORT jlong JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterCreate(
    JNIEnv* env, jobject self, jintArray kinds, jdoubleArray params, jint obs_kind, jboolean has_scale, jdouble scale,
    jint step_mode, jint precision, jint obs_df, jlong n, jint resample_kind, jint dtype, jint device, jlong seed,
    jlong stream_id) {
  cssm_model_desc_t desc; cssm_leaf_t* leaves; jint* k; jdouble* p; cssm_filter_t* f = NULL;
  build_desc(env, kinds, params, obs_kind, has_scale, scale, step_mode, precision, obs_df, &desc, &leaves, &k, &p);
  int rc = cssm_filter_create(&desc, n, resample_kind, dtype, device, (uint64_t)seed, (uint64_t)stream_id, &f);
  free(leaves);
  (*env)->ReleaseIntArrayElements(env, kinds, k, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, params, p, JNI_ABORT);
  check(env, rc);
  return (jlong)(intptr_t)f;
}

JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterSetParams(
    JNIEnv* env, jobject self, jlong h, jintArray kinds, jdoubleArray params, jint obs_kind, jboolean has_scale,
    jdouble scale, jint step_mode, jint precision, jint obs_df) {
  cssm_model_desc_t desc; cssm_leaf_t* leaves; jint* k; jdouble* p;
  build_desc(env, kinds, params, obs_kind, has_scale, scale, step_mode, precision, obs_df, &desc, &leaves, &k, &p);
  int rc = cssm_filter_set_params((cssm_filter_t*)(intptr_t)h, &desc);
  free(leaves);
  (*env)->ReleaseIntArrayElements(env, kinds, k, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, params, p, JNI_ABORT);
  check(env, rc);
}

JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterDestroy(JNIEnv* env, jobject self, jlong h) {
  (void)env; (void)self;
  cssm_filter_destroy((cssm_filter_t*)(intptr_t)h);
}
/* the resident-series form PMMH uses (model/PMMH.scala:71 evaluates the same data at every proposal):
 * filterLoadSeries once, then filterSetParams + filterLlResident per proposal */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterLoadSeries(
    JNIEnv* env, jobject self, jlong h, jdoubleArray t, jdoubleArray y, jbyteArray hasObs) {
  (void)self;
  jsize T = (*env)->GetArrayLength(env, t);
  jdouble* tp = (*env)->GetDoubleArrayElements(env, t, NULL);
  jdouble* yp = (*env)->GetDoubleArrayElements(env, y, NULL);
  jbyte* hp = (*env)->GetByteArrayElements(env, hasObs, NULL);
  int rc = cssm_filter_load_series((cssm_filter_t*)(intptr_t)h, tp, yp, (const uint8_t*)hp, T);
  (*env)->ReleaseDoubleArrayElements(env, t, tp, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, y, yp, JNI_ABORT);
  (*env)->ReleaseByteArrayElements(env, hasObs, hp, JNI_ABORT);
  check(env, rc);
}
JNIEXPORT jdouble JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterLlResident(JNIEnv* env, jobject self, jlong h) {
  (void)self;
  double ll = 0;
  check(env, cssm_filter_ll_resident((cssm_filter_t*)(intptr_t)h, &ll, NULL, NULL));
  return ll;
}
JNIEXPORT jlong JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterSeriesLen(JNIEnv* env, jobject self, jlong h) {
  (void)self;
  int64_t n = 0;
  check(env, cssm_filter_series_len((const cssm_filter_t*)(intptr_t)h, &n));
  return n;
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterSetTieRule(JNIEnv* env, jobject self, jlong h, jint rule) {
  (void)self;
  check(env, cssm_filter_set_tie_rule((cssm_filter_t*)(intptr_t)h, rule));
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterReseed(JNIEnv* env, jobject self, jlong h, jlong seed,
                                                                                   jlong streamId) {
  (void)self;
  check(env, cssm_filter_reseed((cssm_filter_t*)(intptr_t)h, (uint64_t)seed, (uint64_t)streamId));
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterInit(JNIEnv* env, jobject self, jlong h, jdouble t0) {
  check(env, cssm_filter_init((cssm_filter_t*)(intptr_t)h, t0));
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterInitState(JNIEnv* env, jobject self, jlong h,
                                                                                      jdouble t0, jdoubleArray x0) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, x0, NULL);
  int rc = cssm_filter_init_state((cssm_filter_t*)(intptr_t)h, t0, p);
  (*env)->ReleaseDoubleArrayElements(env, x0, p, JNI_ABORT);
  check(env, rc);
}
/* returns ll; ess through essOut[0] */
JNIEXPORT jdouble JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterStep(
    JNIEnv* env, jobject self, jlong h, jdouble t, jboolean has_obs, jdouble y, jintArray essOut) {
  double ll = 0; int32_t ess = 0;
  if (check(env, cssm_filter_step((cssm_filter_t*)(intptr_t)h, t, has_obs, y, &ll, &ess))) return 0;
  jint e = ess;
  (*env)->SetIntArrayRegion(env, essOut, 0, 1, &e);
  return ll;
}
JNIEXPORT jdouble JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterLl(
    JNIEnv* env, jobject self, jlong h, jdoubleArray t, jdoubleArray y, jbyteArray hasObs) {
  jsize T = (*env)->GetArrayLength(env, t);
  jdouble* tp = (*env)->GetDoubleArrayElements(env, t, NULL);
  jdouble* yp = (*env)->GetDoubleArrayElements(env, y, NULL);
  jbyte* hp = (*env)->GetByteArrayElements(env, hasObs, NULL);
  double ll = 0;
  int rc = cssm_filter_ll((cssm_filter_t*)(intptr_t)h, tp, yp, (const uint8_t*)hp, T, &ll);
  (*env)->ReleaseDoubleArrayElements(env, t, tp, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, y, yp, JNI_ABORT);
  (*env)->ReleaseByteArrayElements(env, hasObs, hp, JNI_ABORT);
  check(env, rc);
  return ll;
}
/* filter(): ll returned, statesOut[(T+1)*d] filled */
JNIEXPORT jdouble JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterRun(
    JNIEnv* env, jobject self, jlong h, jdoubleArray t, jdoubleArray y, jbyteArray hasObs, jdoubleArray statesOut) {
  jsize T = (*env)->GetArrayLength(env, t);
  jdouble* tp = (*env)->GetDoubleArrayElements(env, t, NULL);
  jdouble* yp = (*env)->GetDoubleArrayElements(env, y, NULL);
  jbyte* hp = (*env)->GetByteArrayElements(env, hasObs, NULL);
  jdouble* sp = (*env)->GetDoubleArrayElements(env, statesOut, NULL);
  double ll = 0;
  int rc = cssm_filter_run((cssm_filter_t*)(intptr_t)h, tp, yp, (const uint8_t*)hp, T, &ll, sp);
  (*env)->ReleaseDoubleArrayElements(env, t, tp, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, y, yp, JNI_ABORT);
  (*env)->ReleaseByteArrayElements(env, hasObs, hp, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, statesOut, sp, 0);
  check(env, rc);
  return ll;
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterGetParticles(JNIEnv* env, jobject self, jlong h,
                                                                                         jdoubleArray out) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_get_particles((cssm_filter_t*)(intptr_t)h, p);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_resample(
    JNIEnv* env, jobject self, jint kind, jdoubleArray w, jdoubleArray u, jintArray anc, jint device) {
  jsize n = (*env)->GetArrayLength(env, w), nu = (*env)->GetArrayLength(env, u);
  jdouble* wp = (*env)->GetDoubleArrayElements(env, w, NULL);
  jdouble* up = (*env)->GetDoubleArrayElements(env, u, NULL);
  jint* ap = (*env)->GetIntArrayElements(env, anc, NULL);
  int rc = cssm_resample(kind, wp, n, up, nu, (int32_t*)ap, device);
  (*env)->ReleaseDoubleArrayElements(env, w, wp, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, u, up, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, anc, ap, 0);
  check(env, rc);
}

/* cssm_filter_series_mode: 0 auto, 1 three launches per observation, 2 one cooperative launch per llFilter */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterSeriesMode(JNIEnv* env, jobject self, jlong h, jint mode) {
  check(env, cssm_filter_series_mode((cssm_filter_t*)(intptr_t)h, mode));
}
/* ParticleFilter.getIntervals on the device: out = mean[d] | lower[d] | upper[d] | gamma_lower, gamma_upper */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterIntervals(
    JNIEnv* env, jobject self, jlong h, jdouble t, jdouble interval, jint d, jdoubleArray out) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_intervals((cssm_filter_t*)(intptr_t)h, t, interval, p, p + d, p + 2 * d, p + 3 * d);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterSampleOne(JNIEnv* env, jobject self, jlong h,
                                                                                      jdoubleArray out) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_sample_one((cssm_filter_t*)(intptr_t)h, p);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
/* ParticleFilter.getMeanForecast on the device: out = mean[d] | lower[d] | upper[d] | eta mean, lower, upper | obs mean,
 * lower, upper; chain: continue from the previous forecast cloud (SimulateData.forecast) */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterForecast(
    JNIEnv* env, jobject self, jlong h, jdouble t, jdouble interval, jboolean chain, jint d, jdoubleArray out) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_forecast((cssm_filter_t*)(intptr_t)h, t, interval, chain ? 1 : 0, p, p + d, p + 2 * d, p + 3 * d, p + 3 * d + 3);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
/* ParticleFilter.getForecast: the forecast cloud, x[d*N] | gamma[N] | eta[N] | obs[N] */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterForecastCloud(JNIEnv* env, jobject self, jlong h, jint d,
                                                                                          jlong n, jdoubleArray out) {
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_forecast_cloud((cssm_filter_t*)(intptr_t)h, p, p + (size_t)d * n, p + (size_t)(d + 1) * n,
                                      p + (size_t)(d + 2) * n, NULL);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
/* FilterInterpolate: path storage */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterPathsEnable(JNIEnv* env, jobject self, jlong h,
                                                                                        jlong maxSteps) {
  check(env, cssm_filter_paths_enable((cssm_filter_t*)(intptr_t)h, maxSteps));
}
JNIEXPORT jlong JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterPathsLen(JNIEnv* env, jobject self, jlong h) {
  int64_t n = -1;
  check(env, cssm_filter_paths_len((const cssm_filter_t*)(intptr_t)h, &n));
  return n;
}
/* out[idx.length][len + 1][d], oldest state first */
JNIEXPORT void JNICALL Java_com_github_jonnylaw_gpu_CssmNative_00024_filterGetPaths(JNIEnv* env, jobject self, jlong h,
                                                                                     jintArray idx, jdoubleArray out) {
  jsize n = (*env)->GetArrayLength(env, idx);
  jint* ip = (*env)->GetIntArrayElements(env, idx, NULL);
  jdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = cssm_filter_get_paths((cssm_filter_t*)(intptr_t)h, (const int32_t*)ip, n, p);
  (*env)->ReleaseIntArrayElements(env, idx, ip, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, out, p, 0);
  check(env, rc);
}
