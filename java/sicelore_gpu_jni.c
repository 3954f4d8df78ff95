/* JNI glue between com.rw.gpu.Native (java/com/rw/gpu/Native.java) and the C ABI of libsicelore_gpu.so (include/sicelore_gpu.h).
 * Build where a JDK exists:  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/sicelore_gpu_jni.c \
 *                                -Lsicelore-2.1_b200 -lsicelore_gpu -o libsicelore_gpu_jni.so
 * Pure pointer forwarding: direct ByteBuffers are passed by address, Java arrays are pinned for the duration of the call. */
#include <jni.h>
#include <stddef.h>
#include "sicelore_gpu.h"

#define BUF(b) ((b) ? (*env)->GetDirectBufferAddress(env, (b)) : NULL)

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_ctxCreate(JNIEnv *env, jclass cls, jint device, jint nStreams)
{
    (void)env; (void)cls;
    slr_ctx *c = NULL;
    return slr_ctx_create(device, nStreams, &c) == SLR_OK ? (jlong)(size_t)c : 0;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_ctxDestroy(JNIEnv *env, jclass cls, jlong ctx)
{
    (void)env; (void)cls;
    slr_ctx_destroy((slr_ctx *)(size_t)ctx);
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_bcTableCreate(JNIEnv *env, jclass cls, jlong ctx, jlongArray barcodes, jintArray rank)
{
    (void)cls;
    const jsize n = (*env)->GetArrayLength(env, barcodes);
    jlong *b = (*env)->GetLongArrayElements(env, barcodes, NULL);
    jint *r = rank ? (*env)->GetIntArrayElements(env, rank, NULL) : NULL;
    slr_bc_table *t = NULL;
    const int rc = slr_bc_table_create((slr_ctx *)(size_t)ctx, (const uint64_t *)b, (const int32_t *)r, (int64_t)n, 16, &t);
    (*env)->ReleaseLongArrayElements(env, barcodes, b, JNI_ABORT);
    if (r) (*env)->ReleaseIntArrayElements(env, rank, r, JNI_ABORT);
    return rc == SLR_OK ? (jlong)(size_t)t : 0;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_bcTableDestroy(JNIEnv *env, jclass cls, jlong table)
{
    (void)env; (void)cls;
    slr_bc_table_destroy((slr_bc_table *)(size_t)table);
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_bcTableSize(JNIEnv *env, jclass cls, jlong table)
{
    (void)env; (void)cls;
    return (jlong)slr_bc_table_size((const slr_bc_table *)(size_t)table);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcAssign(JNIEnv *env, jclass cls, jlong ctx, jlong table, jint edMax, jint plusMinus,
                                                       jboolean threePrime, jobject slices, jint stride, jint sliceLen, jobject lens,
                                                       jobject anchor, jlong n, jobject out)
{
    (void)cls;
    return slr_bc_assign((slr_ctx *)(size_t)ctx, (const slr_bc_table *)(size_t)table, edMax, plusMinus, threePrime ? 1 : 0,
                         (const uint8_t *)BUF(slices), stride, sliceLen, (const int32_t *)BUF(lens), (const int32_t *)BUF(anchor),
                         (int64_t)n, (slr_bc_result *)BUF(out));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcExact(JNIEnv *env, jclass cls, jlong ctx, jlong table, jboolean threePrime, jobject slices,
                                                      jint stride, jint sliceLen, jobject lens, jobject anchor, jlong n, jobject out)
{
    (void)cls;
    return slr_bc_exact((slr_ctx *)(size_t)ctx, (const slr_bc_table *)(size_t)table, threePrime ? 1 : 0, (const uint8_t *)BUF(slices),
                        stride, sliceLen, (const int32_t *)BUF(lens), (const int32_t *)BUF(anchor), (int64_t)n, (slr_bc_result *)BUF(out));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCountsRead(JNIEnv *env, jclass cls, jlong ctx, jlong table, jlongArray countsOut)
{
    (void)cls;
    jlong *c = (*env)->GetLongArrayElements(env, countsOut, NULL);
    const int rc = slr_bc_counts_read((slr_ctx *)(size_t)ctx, (const slr_bc_table *)(size_t)table, (int64_t *)c);
    (*env)->ReleaseLongArrayElements(env, countsOut, c, 0);
    return rc;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCountsReset(JNIEnv *env, jclass cls, jlong ctx, jlong table)
{
    (void)env; (void)cls;
    return slr_bc_counts_reset((slr_ctx *)(size_t)ctx, (slr_bc_table *)(size_t)table);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCollide(JNIEnv *env, jclass cls, jlong ctx, jlong table, jint edMax, jlongArray barcodes,
                                                        jobject out)
{
    (void)cls;
    const jsize n = (*env)->GetArrayLength(env, barcodes);
    jlong *b = (*env)->GetLongArrayElements(env, barcodes, NULL);
    const int rc = slr_bc_collide((slr_ctx *)(size_t)ctx, (const slr_bc_table *)(size_t)table, edMax, (const uint64_t *)b, (int64_t)n,
                                  (slr_collide_result *)BUF(out));
    (*env)->ReleaseLongArrayElements(env, barcodes, b, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiDist(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                      jobject jobOffsets, jlong nJobs, jobject out, jobject outOffsets)
{
    (void)cls;
    return slr_umi_dist((slr_ctx *)(size_t)ctx, (const uint8_t *)BUF(umis), stride, umiLen, (const int64_t *)BUF(jobOffsets), (int64_t)nJobs,
                        (int32_t *)BUF(out), (const int64_t *)BUF(outOffsets));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiCluster(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                         jobject jobOffsets, jlong nJobs, jint ed, jobject member, jobject rank, jobject out,
                                                         jobject outOffsets, jobject rec)
{
    (void)cls;
    return slr_umi_cluster((slr_ctx *)(size_t)ctx, (const uint8_t *)BUF(umis), stride, umiLen, (const int64_t *)BUF(jobOffsets), (int64_t)nJobs, ed,
                           member ? (const uint8_t *)BUF(member) : NULL, rank ? (const int32_t *)BUF(rank) : NULL,
                           out ? (int32_t *)BUF(out) : NULL, outOffsets ? (const int64_t *)BUF(outOffsets) : NULL, (slr_umi_cluster_rec *)BUF(rec));
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_umiSessionCreate(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                                jobject jobOffsets, jlong nJobs)
{
    (void)cls;
    slr_umi_session *s = NULL;
    const int rc = slr_umi_session_create((slr_ctx *)(size_t)ctx, (const uint8_t *)BUF(umis), stride, umiLen, (const int64_t *)BUF(jobOffsets),
                                          (int64_t)nJobs, &s);
    return rc == SLR_OK ? (jlong)(size_t)s : 0;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiSessionCluster(JNIEnv *env, jclass cls, jlong session, jint ed, jobject member, jobject rank,
                                                                jobject rec)
{
    (void)cls;
    return slr_umi_session_cluster((slr_umi_session *)(size_t)session, ed, member ? (const uint8_t *)BUF(member) : NULL,
                                   rank ? (const int32_t *)BUF(rank) : NULL, (slr_umi_cluster_rec *)BUF(rec));
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_umiSessionCells(JNIEnv *env, jclass cls, jlong session)
{
    (void)env; (void)cls;
    return (jlong)slr_umi_session_cells((const slr_umi_session *)(size_t)session);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiSessionMatrices(JNIEnv *env, jclass cls, jlong session, jobject out, jlong nCells)
{
    (void)cls;
    return slr_umi_session_matrices((slr_umi_session *)(size_t)session, (int32_t *)BUF(out), (int64_t)nCells);
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_umiSessionDestroy(JNIEnv *env, jclass cls, jlong session)
{
    (void)env; (void)cls;
    slr_umi_session_destroy((slr_umi_session *)(size_t)session);
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_guidedSetsCreate(JNIEnv *env, jclass cls, jlong ctx, jlongArray groupKeys, jlongArray groupOffsets,
                                                                jlongArray allKeys, jint allEd, jlongArray emptyKeys, jint emptyEd,
                                                                jboolean bcFlavour, jint seqLen)
{
    (void)cls;
    const jsize ng = (*env)->GetArrayLength(env, groupOffsets) - 1;
    jlong *gk = (*env)->GetLongArrayElements(env, groupKeys, NULL);
    jlong *go = (*env)->GetLongArrayElements(env, groupOffsets, NULL);
    jlong *ak = allKeys ? (*env)->GetLongArrayElements(env, allKeys, NULL) : NULL;
    jlong *ek = emptyKeys ? (*env)->GetLongArrayElements(env, emptyKeys, NULL) : NULL;
    slr_guided_sets *s = NULL;
    const int rc = slr_guided_sets_create((slr_ctx *)(size_t)ctx, (const uint64_t *)gk, (const int64_t *)go, (int64_t)ng, (const uint64_t *)ak,
                                          ak ? (int64_t)(*env)->GetArrayLength(env, allKeys) : 0, allEd, (const uint64_t *)ek,
                                          ek ? (int64_t)(*env)->GetArrayLength(env, emptyKeys) : 0, emptyEd, bcFlavour ? 1 : 0, seqLen, &s);
    (*env)->ReleaseLongArrayElements(env, groupKeys, gk, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, groupOffsets, go, JNI_ABORT);
    if (ak) (*env)->ReleaseLongArrayElements(env, allKeys, ak, JNI_ABORT);
    if (ek) (*env)->ReleaseLongArrayElements(env, emptyKeys, ek, JNI_ABORT);
    return rc == SLR_OK ? (jlong)(size_t)s : 0;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_guidedSetsDestroy(JNIEnv *env, jclass cls, jlong sets)
{
    (void)env; (void)cls;
    slr_guided_sets_destroy((slr_guided_sets *)(size_t)sets);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_guidedMatch(JNIEnv *env, jclass cls, jlong ctx, jlong sets, jint plusMinus, jint postLen, jint bailout,
                                                          jobject slices, jint stride, jint sliceLen, jobject anchor, jobject groupId, jobject ed,
                                                          jlong n, jobject out, jobject rawOut, jint rawCap)
{
    (void)cls;
    return slr_guided_match((slr_ctx *)(size_t)ctx, (const slr_guided_sets *)(size_t)sets, plusMinus, postLen, bailout, (const uint8_t *)BUF(slices),
                            stride, sliceLen, (const int32_t *)BUF(anchor), (const int32_t *)BUF(groupId), (const int32_t *)BUF(ed), (int64_t)n,
                            (slr_guided_result *)BUF(out), rawOut ? (slr_guided_hit *)BUF(rawOut) : NULL, rawCap);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_dynMaxEd(JNIEnv *env, jclass cls, jlongArray maxCandidates, jint count, jint plusMinus, jint cap)
{
    (void)cls;
    const jsize n = (*env)->GetArrayLength(env, maxCandidates);
    jlong *m = (*env)->GetLongArrayElements(env, maxCandidates, NULL);
    const int r = slr_dyn_max_ed((const int64_t *)m, (int)n, count, plusMinus, cap);
    (*env)->ReleaseLongArrayElements(env, maxCandidates, m, JNI_ABORT);
    return r;
}

JNIEXPORT jstring JNICALL Java_com_rw_gpu_Native_lastError(JNIEnv *env, jclass cls)
{
    (void)cls;
    return (*env)->NewStringUTF(env, slr_last_error());
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_abiVersion(JNIEnv *env, jclass cls)
{
    (void)env; (void)cls;
    return slr_abi_version();
}
