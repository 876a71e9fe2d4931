/*
 * rafting_jni.c — JNI glue between io.lubricant.consensus.raft.gpu.NativeEngine and the C ABI
 * (include/rafting_b200.h).  It only marshals DirectByteBuffer addresses; no protocol logic.
 *
 * NOT COMPILED IN THIS IMAGE: there is no JDK (no jni.h).  The whole file is guarded so that the
 * build never depends on it; on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       rafting_b200/csrc/jni/rafting_jni.c -Lrafting_b200 -lrafting_b200 -o librafting_jni.so
 * The Java side is shown in INTEGRATION.md.
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#define RAFTING_HAVE_JNI 1
#endif
#endif

#ifdef RAFTING_HAVE_JNI
#include <jni.h>
#include <string.h>

#include "../../../include/rafting_b200.h"

#define FN(name) Java_io_lubricant_consensus_raft_gpu_NativeEngine_##name

static void throw_status(JNIEnv* env, int rc) {
    /* status codes map 1:1 onto the reference's exception types */
    const char* cls = rc == RAFTING_E_INVAL ? "java/lang/IllegalArgumentException"
                    : rc == RAFTING_E_CLOSED ? "java/lang/IllegalStateException"
                    : "io/lubricant/consensus/raft/support/RaftException";
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, rafting_last_error());
}

JNIEXPORT jlong JNICALL FN(create)(JNIEnv* env, jclass k, jobject cfgBuf) {
    rafting_engine_t* e = NULL;
    const rafting_cfg_t* cfg = (const rafting_cfg_t*)(*env)->GetDirectBufferAddress(env, cfgBuf);
    int rc = rafting_engine_create(cfg, &e);
    if (rc) { throw_status(env, rc); return 0; }
    return (jlong)(intptr_t)e;
}
JNIEXPORT void JNICALL FN(destroy)(JNIEnv* env, jclass k, jlong h) { rafting_engine_destroy((rafting_engine_t*)(intptr_t)h); }

JNIEXPORT void JNICALL FN(groupOpen)(JNIEnv* env, jclass k, jlong h, jint gid, jobject initBuf) {
    int rc = rafting_group_open((rafting_engine_t*)(intptr_t)h, (uint32_t)gid,
                                (const rafting_group_init_t*)(*env)->GetDirectBufferAddress(env, initBuf));
    if (rc) throw_status(env, rc);
}
JNIEXPORT void JNICALL FN(groupClose)(JNIEnv* env, jclass k, jlong h, jint gid) {
    int rc = rafting_group_close((rafting_engine_t*)(intptr_t)h, (uint32_t)gid);
    if (rc) throw_status(env, rc);
}

/* lease: fills `leaseBuf` (a direct buffer of sizeof(rafting_lease_t)) and returns the pinned columns as
   DirectByteBuffers the pump thread writes events into — see NativeEngine.Lease in INTEGRATION.md */
JNIEXPORT void JNICALL FN(lease)(JNIEnv* env, jclass k, jlong h, jint rows, jint nActive, jint entCount, jint flags, jobject leaseBuf) {
    rafting_lease_t* L = (rafting_lease_t*)(*env)->GetDirectBufferAddress(env, leaseBuf);
    int rc = rafting_lease_ex((rafting_engine_t*)(intptr_t)h, (uint32_t)rows, (uint32_t)nActive, (uint32_t)entCount, (uint32_t)flags, L);
    if (rc) throw_status(env, rc);
}
JNIEXPORT jobject JNICALL FN(wrap)(JNIEnv* env, jclass k, jlong addr, jlong bytes) {
    return (*env)->NewDirectByteBuffer(env, (void*)(intptr_t)addr, bytes);
}
JNIEXPORT void JNICALL FN(stepBegin)(JNIEnv* env, jclass k, jlong h, jobject leaseBuf) {
    int rc = rafting_step_begin((rafting_engine_t*)(intptr_t)h, (rafting_lease_t*)(*env)->GetDirectBufferAddress(env, leaseBuf));
    if (rc) throw_status(env, rc);
}
JNIEXPORT void JNICALL FN(stepWait)(JNIEnv* env, jclass k, jlong h, jobject leaseBuf) {
    int rc = rafting_step_wait((rafting_engine_t*)(intptr_t)h, (rafting_lease_t*)(*env)->GetDirectBufferAddress(env, leaseBuf));
    if (rc) throw_status(env, rc);
}
JNIEXPORT void JNICALL FN(stateExport)(JNIEnv* env, jclass k, jlong h, jint gid, jobject outBuf) {
    int rc = rafting_state_export((rafting_engine_t*)(intptr_t)h, (uint32_t)gid,
                                  (rafting_group_state_t*)(*env)->GetDirectBufferAddress(env, outBuf));
    if (rc) throw_status(env, rc);
}
JNIEXPORT jlong JNICALL FN(logTerm)(JNIEnv* env, jclass k, jlong h, jint gid, jlong index) {
    int64_t t = -1;
    int rc = rafting_log_term((rafting_engine_t*)(intptr_t)h, (uint32_t)gid, index, &t);
    if (rc) throw_status(env, rc);
    return t;
}
JNIEXPORT void JNICALL FN(allgatherCommit)(JNIEnv* env, jclass k, jlong h, jobject hostOut) {
    int rc = rafting_allgather_commit((rafting_engine_t*)(intptr_t)h,
                                      hostOut ? (int64_t*)(*env)->GetDirectBufferAddress(env, hostOut) : NULL, NULL);
    if (rc) throw_status(env, rc);
}
#endif /* RAFTING_HAVE_JNI */
