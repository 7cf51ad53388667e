/* jni_harness.c -- drives the Java_... entry points of jvm/cssm_jni.c through a FAKE JNIEnv (jvm/jni_stub/jni.h): arrays
 * are plain C buffers with a length header, ThrowNew records the message.  Test infrastructure (tests/test_jni_shim.py).
 *
 *   jni_harness errors      no GPU needed: every call is rejected before the device is touched; prints what was thrown
 *   jni_harness filter      needs a GPU: create -> loadSeries -> llResident (twice, reseeded: same bits) -> setParams ->
 *                           llResident -> filterLl -> step API; prints the numbers as JSON for the test to compare
 */
#include <jni.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { jsize len; int elem; double pad; } arr_hdr;   /* the payload follows the header */
static char g_thrown[1024];
static int g_throws = 0;

static jarray mk(int elem, jsize len, const void* init) {
  arr_hdr* h = (arr_hdr*)calloc(1, sizeof(arr_hdr) + (size_t)elem * (size_t)(len > 0 ? len : 1));
  h->len = len; h->elem = elem;
  if (init) memcpy(h + 1, init, (size_t)elem * (size_t)len);
  return (jarray)h;
}
static void* payload(jarray a) { return (void*)((arr_hdr*)a + 1); }

static jclass f_FindClass(JNIEnv* env, const char* name) { (void)env; return (jclass)name; }
static jint f_ThrowNew(JNIEnv* env, jclass c, const char* msg) {
  (void)env; (void)c; ++g_throws; snprintf(g_thrown, sizeof g_thrown, "%s", msg ? msg : ""); return 0;
}
static jsize f_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return ((arr_hdr*)a)->len; }
static jint* f_GetInt(JNIEnv* env, jintArray a, jboolean* c) { (void)env; if (c) *c = 0; return (jint*)payload(a); }
static jdouble* f_GetDouble(JNIEnv* env, jdoubleArray a, jboolean* c) { (void)env; if (c) *c = 0; return (jdouble*)payload(a); }
static jbyte* f_GetByte(JNIEnv* env, jbyteArray a, jboolean* c) { (void)env; if (c) *c = 0; return (jbyte*)payload(a); }
static void f_RelInt(JNIEnv* env, jintArray a, jint* e, jint m) { (void)env; (void)a; (void)e; (void)m; }
static void f_RelDouble(JNIEnv* env, jdoubleArray a, jdouble* e, jint m) { (void)env; (void)a; (void)e; (void)m; }
static void f_RelByte(JNIEnv* env, jbyteArray a, jbyte* e, jint m) { (void)env; (void)a; (void)e; (void)m; }
static void f_SetIntRegion(JNIEnv* env, jintArray a, jsize s, jsize n, const jint* buf) {
  (void)env; memcpy((jint*)payload(a) + s, buf, sizeof(jint) * (size_t)n);
}
static const struct JNINativeInterface_ g_table = {f_FindClass, f_ThrowNew, f_GetArrayLength, f_GetInt, f_GetDouble, f_GetByte,
                                                   f_RelInt, f_RelDouble, f_RelByte, f_SetIntRegion};

/* the shim's entry points (CssmNative is a Scala object: the JNI class name is CssmNative$ -> _00024) */
#define N(x) Java_com_github_jonnylaw_gpu_CssmNative_00024_##x
jlong N(filterCreate)(JNIEnv*, jobject, jintArray, jdoubleArray, jint, jboolean, jdouble, jint, jint, jint, jlong, jint, jint, jint, jlong, jlong);
void N(filterSetParams)(JNIEnv*, jobject, jlong, jintArray, jdoubleArray, jint, jboolean, jdouble, jint, jint, jint);
void N(filterDestroy)(JNIEnv*, jobject, jlong);
void N(filterInit)(JNIEnv*, jobject, jlong, jdouble);
jdouble N(filterStep)(JNIEnv*, jobject, jlong, jdouble, jboolean, jdouble, jintArray);
jdouble N(filterLl)(JNIEnv*, jobject, jlong, jdoubleArray, jdoubleArray, jbyteArray);
void N(filterLoadSeries)(JNIEnv*, jobject, jlong, jdoubleArray, jdoubleArray, jbyteArray);
jdouble N(filterLlResident)(JNIEnv*, jobject, jlong);
jlong N(filterSeriesLen)(JNIEnv*, jobject, jlong);
void N(filterReseed)(JNIEnv*, jobject, jlong, jlong, jlong);
void N(filterSeriesMode)(JNIEnv*, jobject, jlong, jint);
void N(filterSampleOne)(JNIEnv*, jobject, jlong, jdoubleArray);
void N(filterGetParticles)(JNIEnv*, jobject, jlong, jdoubleArray);
void N(resample)(JNIEnv*, jobject, jint, jdoubleArray, jdoubleArray, jintArray, jint);

/* Model.poisson(Sde.ouProcess(1)) with the reference's example values (examples/Simulation.scala:16), flattened as
 * GpuDesc does: kinds = (sde_kind, dim, f_kind, period, harmonics), params = m0 | c0 | phi | mu | sigma */
static jintArray kinds_c1(void) { const jint k[5] = {2, 1, 0, 0, 0}; return mk(sizeof(jint), 5, k); }
static jdoubleArray params_c1(double sigma) {
  const jdouble p[5] = {1.0, 0.5, 0.6340970762609854, 1.5, sigma};
  return mk(sizeof(jdouble), 5, p);
}

int main(int argc, char** argv) {
  const struct JNINativeInterface_* tbl = &g_table;
  JNIEnv* env = &tbl;
  const char* mode = argc > 1 ? argv[1] : "errors";
  if (strcmp(mode, "errors") == 0) {
    /* 0 particles: rejected with the library's message, rethrown as RuntimeException */
    jlong h = N(filterCreate)(env, NULL, kinds_c1(), params_c1(0.05), 0, 0, 0.0, 0, 0, 0, 0, 0, 0, 0, 1, 0);
    printf("create(n=0): handle %lld throws %d msg \"%s\"\n", (long long)h, g_throws, g_thrown);
    N(filterSeriesMode)(env, NULL, 0, 7);
    printf("seriesMode(null handle): throws %d msg \"%s\"\n", g_throws, g_thrown);
    jdouble ll = N(filterLlResident)(env, NULL, 0);
    printf("llResident(null handle): ll %g throws %d msg \"%s\"\n", ll, g_throws, g_thrown);
    N(filterDestroy)(env, NULL, 0);
    printf("destroy(null handle): throws %d\n", g_throws);
    return g_throws == 3 ? 0 : 1;
  }
  /* ---- filter: needs a GPU ---- */
  enum { T = 40, NP = 4096 };
  jdouble t[T], y[T];
  jbyte ho[T];
  for (int s = 0; s < T; ++s) { t[s] = 0.1 * s; y[s] = (double)((s * 7 + 3) % 6); ho[s] = (s == 5) ? 0 : 1; }
  jdoubleArray ta = mk(sizeof(jdouble), T, t), ya = mk(sizeof(jdouble), T, y);
  jbyteArray ha = mk(sizeof(jbyte), T, ho);
  jlong h = N(filterCreate)(env, NULL, kinds_c1(), params_c1(0.05), 0, 0, 0.0, 0, 0, 0, NP, 0, 1 /* F64 */, 0, 11, 0);
  if (g_throws || !h) { printf("{\"error\": \"%s\"}\n", g_thrown); return 2; }
  N(filterLoadSeries)(env, NULL, h, ta, ya, ha);
  const jlong len = N(filterSeriesLen)(env, NULL, h);
  const jdouble ll1 = N(filterLlResident)(env, NULL, h);
  N(filterReseed)(env, NULL, h, 11, 0);
  const jdouble ll1b = N(filterLlResident)(env, NULL, h);      /* same seed: same bits */
  N(filterSetParams)(env, NULL, h, kinds_c1(), params_c1(0.2), 0, 0, 0.0, 0, 0, 0);
  N(filterReseed)(env, NULL, h, 11, 0);
  const jdouble ll2 = N(filterLlResident)(env, NULL, h);       /* other parameters */
  N(filterReseed)(env, NULL, h, 11, 0);
  const jdouble ll3 = N(filterLl)(env, NULL, h, ta, ya, ha);   /* host buffers in, same parameters and seed as ll2 */
  jdoubleArray one = mk(sizeof(jdouble), 1, NULL), cloud = mk(sizeof(jdouble), NP, NULL);
  N(filterSampleOne)(env, NULL, h, one);
  N(filterGetParticles)(env, NULL, h, cloud);
  int member = 0;
  for (int i = 0; i < NP; ++i) member |= ((jdouble*)payload(cloud))[i] == ((jdouble*)payload(one))[0];
  /* the stepping API: initialiseState + stepFilter */
  N(filterReseed)(env, NULL, h, 11, 0);
  N(filterInit)(env, NULL, h, t[0]);
  jintArray ess = mk(sizeof(jint), 1, NULL);
  jdouble ll4 = 0;
  for (int s = 0; s < T; ++s) ll4 = N(filterStep)(env, NULL, h, t[s], ho[s] != 0, y[s], ess);
  /* Resample[A] */
  const jdouble w[4] = {0.1, 0.2, 0.3, 0.4}, u[1] = {0.5};
  jintArray anc = mk(sizeof(jint), 4, NULL);
  N(resample)(env, NULL, 0, mk(sizeof(jdouble), 4, w), mk(sizeof(jdouble), 1, u), anc, 0);
  const jint* a = (const jint*)payload(anc);
  N(filterDestroy)(env, NULL, h);
  printf("{\"series_len\": %lld, \"ll1\": %.17g, \"ll1b\": %.17g, \"ll2\": %.17g, \"ll3\": %.17g, \"ll4\": %.17g, \"ess\": %d, "
         "\"sample_is_member\": %d, \"anc\": [%d, %d, %d, %d], \"throws\": %d, \"thrown\": \"%s\"}\n",
         (long long)len, ll1, ll1b, ll2, ll3, ll4, ((jint*)payload(ess))[0], member, a[0], a[1], a[2], a[3], g_throws, g_thrown);
  return g_throws ? 3 : 0;
}
