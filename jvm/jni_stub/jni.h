/* jni.h -- a STAND-IN for the JDK's header, for images without a JDK (this one has none).
 *
 * Only what jvm/cssm_jni.c uses: the primitive and array typedefs, JNIEXPORT / JNICALL, and a JNIEnv whose
 * function table names the handful of functions the shim calls.  The member NAMES and SIGNATURES are the JDK's, so
 * the shim compiles unchanged against the real <jni.h> (-I$JAVA_HOME/include -I$JAVA_HOME/include/linux); the table
 * LAYOUT is not the JDK's (the real JNINativeInterface_ has 229 slots in a fixed order), so a library compiled against
 * this stand-in must never be loaded by a JVM.  It exists so that `gcc -Wall -Werror` type-checks the shim and
 * tests/jni_harness.c can drive the Java_... entry points through a fake JNIEnv (tests/test_jni_shim.py).
 */
#ifndef CSSM_JNI_STUB_H
#define CSSM_JNI_STUB_H
#define CSSM_JNI_STUB 1
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
typedef jarray jbyteArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
  jdouble* (*GetDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jboolean* isCopy);
  jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
  void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
  void (*ReleaseDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jdouble* elems, jint mode);
  void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
  void (*SetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, const jint* buf);
};
#endif
