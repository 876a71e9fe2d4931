/* Minimal stand-in for <jni.h>: just the types and the JNINativeInterface_ slots rafting_jni.c uses, so that the glue can be
   syntax- and type-checked against the real C-ABI headers on a box without a JDK (tests/test_jni_glue_cpu.py). */
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef unsigned char jboolean; typedef int32_t jsize;
typedef void* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jlongArray;
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_; typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*); jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject); jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*); void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jobject); jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*);
  void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
};
