/* A stand-in JNIEnv for tests/test_jni_glue_cpu.py: enough of the function table (tests/jni_stub/jni.h) to EXECUTE the natives
 * of rafting_b200/csrc/jni/rafting_jni.c on a box without a JDK.  Conventions of this fake VM:
 *   direct ByteBuffer  = fake_buf_t*  {addr, cap}
 *   String             = const char*  (NUL-terminated UTF-8)
 *   long[]             = fake_longs_t* {n, p}
 *   class              = const char*  (the name FindClass was asked for)
 * A thrown exception is recorded (class, message, count) instead of unwinding; the test reads it back. */
#include <stdlib.h>
#include <string.h>

#include "jni.h"

typedef struct { void* addr; int64_t cap; } fake_buf_t;
typedef struct { int32_t n; int64_t* p; } fake_longs_t;

static char g_class[160], g_msg[1024], g_found[160];
static int g_throws, g_string_gets, g_string_releases, g_array_gets, g_array_releases, g_wraps;

static jclass f_FindClass(JNIEnv* e, const char* name) { (void)e; strncpy(g_found, name, sizeof g_found - 1); return (jclass)g_found; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* msg) {
    (void)e; g_throws++;
    strncpy(g_class, c ? (const char*)c : "", sizeof g_class - 1);
    strncpy(g_msg, msg ? msg : "", sizeof g_msg - 1);
    return 0;
}
static void* f_GetDirectBufferAddress(JNIEnv* e, jobject b) { (void)e; return b ? ((fake_buf_t*)b)->addr : NULL; }
static jlong f_GetDirectBufferCapacity(JNIEnv* e, jobject b) { (void)e; return b ? ((fake_buf_t*)b)->cap : -1; }
static jobject f_NewDirectByteBuffer(JNIEnv* e, void* addr, jlong cap) {
    (void)e; fake_buf_t* b = (fake_buf_t*)malloc(sizeof *b); b->addr = addr; b->cap = cap; g_wraps++; return b;
}
static const char* f_GetStringUTFChars(JNIEnv* e, jstring s, jboolean* copy) { (void)e; if (copy) *copy = 0; g_string_gets++; return (const char*)s; }
static void f_ReleaseStringUTFChars(JNIEnv* e, jstring s, const char* c) { (void)e; (void)s; (void)c; g_string_releases++; }
static jsize f_GetArrayLength(JNIEnv* e, jobject a) { (void)e; return ((fake_longs_t*)a)->n; }
static jlong* f_GetLongArrayElements(JNIEnv* e, jlongArray a, jboolean* copy) { (void)e; if (copy) *copy = 0; g_array_gets++; return (jlong*)((fake_longs_t*)a)->p; }
static void f_ReleaseLongArrayElements(JNIEnv* e, jlongArray a, jlong* p, jint mode) { (void)e; (void)a; (void)p; (void)mode; g_array_releases++; }

static const struct JNINativeInterface_ g_table = {
    f_FindClass, f_ThrowNew, f_GetDirectBufferAddress, f_GetDirectBufferCapacity, f_NewDirectByteBuffer,
    f_GetStringUTFChars, f_ReleaseStringUTFChars, f_GetArrayLength, f_GetLongArrayElements, f_ReleaseLongArrayElements,
};
static JNIEnv g_env = &g_table;

JNIEnv* fake_env(void) { return &g_env; }
void fake_reset(void) { g_class[0] = g_msg[0] = g_found[0] = 0; g_throws = g_string_gets = g_string_releases = g_array_gets = g_array_releases = g_wraps = 0; }
int fake_throws(void) { return g_throws; }
const char* fake_thrown_class(void) { return g_class; }
const char* fake_thrown_message(void) { return g_msg; }
int fake_balance(void) { return (g_string_gets - g_string_releases) + (g_array_gets - g_array_releases); }   /* 0 = every Get was Released */
void fake_free_buf(void* b) { free(b); }
