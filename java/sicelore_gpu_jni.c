/* JNI glue between com.rw.gpu.Native (java/com/rw/gpu/Native.java) and the C ABI of libsicelore_gpu.so (include/sicelore_gpu.h).
 * Build where a JDK exists:  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/sicelore_gpu_jni.c \
 *                                -Lsicelore-2.1_b200 -lsicelore_gpu -o libsicelore_gpu_jni.so
 * Pointer forwarding with size checks: direct ByteBuffers are passed by address after their capacity has been compared with the bytes the
 * call will touch (n * stride, n * record size, (nJobs + 1) * 8 ...), Java arrays are pinned for the duration of the call and their lengths
 * checked the same way.  A mismatch returns SLR_E_INVALID (-1) before anything is read or written: a wrong n from the Java side must not
 * corrupt the JVM heap. */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "sicelore_gpu.h"
#include "sicelore_host.h"

#define H(T, x) ((T *)(size_t)(x))

/* address of a direct buffer that must hold at least `need` bytes; *ok is cleared when it does not (or when a required buffer is null) */
static void *buf(JNIEnv *env, jobject b, int64_t need, int required, int *ok)
{
    if (!b) {
        if (required && need > 0) *ok = 0;
        return NULL;
    }
    void *p = (*env)->GetDirectBufferAddress(env, b);
    const jlong cap = (*env)->GetDirectBufferCapacity(env, b);
    if (!p || need < 0 || cap < (jlong)need) { *ok = 0; return NULL; }
    return p;
}
/* total reads of a CSR job list: jobOffsets[nJobs] (the buffer is checked for (nJobs + 1) * 8 bytes first) */
static int64_t csr_reads(JNIEnv *env, jobject jobOffsets, jlong nJobs, int *ok)
{
    if (nJobs < 0) { *ok = 0; return 0; }
    const int64_t *o = (const int64_t *)buf(env, jobOffsets, (int64_t)(nJobs + 1) * 8, 1, ok);
    if (!*ok || !o) return 0;
    if (o[nJobs] < o[0]) { *ok = 0; return 0; }
    return o[nJobs];                                        /* the library addresses reads 0 .. jobOffsets[nJobs] - 1 of `umis` */
}
static int64_t csr_cells(const int64_t *o, jlong nJobs)
{
    int64_t c = 0;
    for (jlong j = 0; j < nJobs; j++) { const int64_t n = o[j + 1] - o[j]; c += n * n; }
    return c;
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_ctxCreate(JNIEnv *env, jclass cls, jint device, jint nStreams)
{
    (void)env; (void)cls;
    slr_ctx *c = NULL;
    return slr_ctx_create(device, nStreams, &c) == SLR_OK ? (jlong)(size_t)c : 0;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_ctxDestroy(JNIEnv *env, jclass cls, jlong ctx)
{
    (void)env; (void)cls;
    slr_ctx_destroy(H(slr_ctx, ctx));
}

static jlong table_create(JNIEnv *env, jlong owner, int multi, jlongArray barcodes, jintArray rank)
{
    if (!owner || !barcodes) return 0;
    const jsize n = (*env)->GetArrayLength(env, barcodes);
    if (rank && (*env)->GetArrayLength(env, rank) < n) return 0;          /* a rank[] shorter than barcodes[] would be read past its end */
    jlong *b = (*env)->GetLongArrayElements(env, barcodes, NULL);
    jint *r = rank ? (*env)->GetIntArrayElements(env, rank, NULL) : NULL;
    jlong res = 0;
    if (b && (r || !rank)) {
        if (multi) {
            slr_multi_table *t = NULL;
            if (slr_multi_bc_table_create(H(slr_multi, owner), (const uint64_t *)b, (const int32_t *)r, (int64_t)n, 16, &t) == SLR_OK) res = (jlong)(size_t)t;
        } else {
            slr_bc_table *t = NULL;
            if (slr_bc_table_create(H(slr_ctx, owner), (const uint64_t *)b, (const int32_t *)r, (int64_t)n, 16, &t) == SLR_OK) res = (jlong)(size_t)t;
        }
    }
    if (b) (*env)->ReleaseLongArrayElements(env, barcodes, b, JNI_ABORT);
    if (r) (*env)->ReleaseIntArrayElements(env, rank, r, JNI_ABORT);
    return res;
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_bcTableCreate(JNIEnv *env, jclass cls, jlong ctx, jlongArray barcodes, jintArray rank)
{
    (void)cls;
    return table_create(env, ctx, 0, barcodes, rank);
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_bcTableDestroy(JNIEnv *env, jclass cls, jlong table)
{
    (void)env; (void)cls;
    slr_bc_table_destroy(H(slr_bc_table, table));
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_bcTableSize(JNIEnv *env, jclass cls, jlong table)
{
    (void)env; (void)cls;
    return (jlong)slr_bc_table_size(H(const slr_bc_table, table));
}

/* the buffers of one barcode batch, checked against n */
typedef struct { const uint8_t *slices; const int32_t *lens, *anchor; slr_bc_result *out; } bc_bufs;
static int bc_check(JNIEnv *env, jobject slices, jint stride, jobject lens, jobject anchor, jlong n, jobject out, bc_bufs *B)
{
    int ok = n >= 0 && stride > 0;
    B->slices = (const uint8_t *)buf(env, slices, (int64_t)n * stride, 1, &ok);
    B->lens = (const int32_t *)buf(env, lens, (int64_t)n * 4, 0, &ok);
    B->anchor = (const int32_t *)buf(env, anchor, (int64_t)n * 4, 1, &ok);
    B->out = (slr_bc_result *)buf(env, out, (int64_t)n * (int64_t)sizeof(slr_bc_result), 1, &ok);
    return ok;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcAssign(JNIEnv *env, jclass cls, jlong ctx, jlong table, jint edMax, jint plusMinus,
                                                       jboolean threePrime, jobject slices, jint stride, jint sliceLen, jobject lens,
                                                       jobject anchor, jlong n, jobject out)
{
    (void)cls;
    bc_bufs B;
    if (!bc_check(env, slices, stride, lens, anchor, n, out, &B)) return SLR_E_INVALID;
    return slr_bc_assign(H(slr_ctx, ctx), H(const slr_bc_table, table), edMax, plusMinus, threePrime ? 1 : 0, B.slices, stride, sliceLen, B.lens,
                         B.anchor, (int64_t)n, B.out);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcExact(JNIEnv *env, jclass cls, jlong ctx, jlong table, jboolean threePrime, jobject slices,
                                                      jint stride, jint sliceLen, jobject lens, jobject anchor, jlong n, jobject out)
{
    (void)cls;
    bc_bufs B;
    if (!bc_check(env, slices, stride, lens, anchor, n, out, &B)) return SLR_E_INVALID;
    return slr_bc_exact(H(slr_ctx, ctx), H(const slr_bc_table, table), threePrime ? 1 : 0, B.slices, stride, sliceLen, B.lens, B.anchor, (int64_t)n,
                        B.out);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCountsRead(JNIEnv *env, jclass cls, jlong ctx, jlong table, jlongArray countsOut)
{
    (void)cls;
    if (!ctx || !table || !countsOut) return SLR_E_INVALID;
    /* the table holds one counter triple per INPUT barcode (duplicates included) */
    int64_t *dc = NULL, ne = 0;
    if (slr_bc_counts_device(H(const slr_bc_table, table), &dc, &ne) != SLR_OK) return SLR_E_INVALID;
    if ((int64_t)(*env)->GetArrayLength(env, countsOut) < ne) return SLR_E_INVALID;
    jlong *c = (*env)->GetLongArrayElements(env, countsOut, NULL);
    if (!c) return SLR_E_NOMEM;
    const int rc = slr_bc_counts_read(H(slr_ctx, ctx), H(const slr_bc_table, table), (int64_t *)c);
    (*env)->ReleaseLongArrayElements(env, countsOut, c, 0);
    return rc;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCountsReset(JNIEnv *env, jclass cls, jlong ctx, jlong table)
{
    (void)env; (void)cls;
    return slr_bc_counts_reset(H(slr_ctx, ctx), H(slr_bc_table, table));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_bcCollide(JNIEnv *env, jclass cls, jlong ctx, jlong table, jint edMax, jlongArray barcodes,
                                                        jobject out)
{
    (void)cls;
    if (!barcodes) return SLR_E_INVALID;
    const jsize n = (*env)->GetArrayLength(env, barcodes);
    int ok = 1;
    slr_collide_result *o = (slr_collide_result *)buf(env, out, (int64_t)n * (int64_t)sizeof(slr_collide_result), 1, &ok);
    if (!ok) return SLR_E_INVALID;
    jlong *b = (*env)->GetLongArrayElements(env, barcodes, NULL);
    if (!b) return SLR_E_NOMEM;
    const int rc = slr_bc_collide(H(slr_ctx, ctx), H(const slr_bc_table, table), edMax, (const uint64_t *)b, (int64_t)n, o);
    (*env)->ReleaseLongArrayElements(env, barcodes, b, JNI_ABORT);
    return rc;
}

/* the buffers of one UMI batch */
typedef struct { const uint8_t *umis; const int64_t *joff; int64_t reads; } umi_bufs;
static int umi_check(JNIEnv *env, jobject umis, jint stride, jobject jobOffsets, jlong nJobs, umi_bufs *U)
{
    int ok = stride > 0;
    U->reads = csr_reads(env, jobOffsets, nJobs, &ok);
    U->joff = ok ? (const int64_t *)(*env)->GetDirectBufferAddress(env, jobOffsets) : NULL;
    U->umis = (const uint8_t *)buf(env, umis, U->reads * stride, 1, &ok);
    return ok;
}
/* the optional matrix output: outOffsets (nJobs + 1 longs) and out (outOffsets[nJobs] ints, at least the cells of the jobs) */
static int mat_check(JNIEnv *env, const umi_bufs *U, jlong nJobs, jobject out, jobject outOffsets, int required, int32_t **o, const int64_t **oo)
{
    int ok = 1;
    *o = NULL; *oo = NULL;
    if (!out && !required) return 1;
    *oo = (const int64_t *)buf(env, outOffsets, (int64_t)(nJobs + 1) * 8, 1, &ok);
    if (!ok) return 0;
    int64_t need = csr_cells(U->joff, nJobs);
    if ((*oo)[nJobs] > need) need = (*oo)[nJobs];
    *o = (int32_t *)buf(env, out, need * 4, 1, &ok);
    return ok;
}
/* int[5] or int[6] clustering parameters (the sixth, `deep`, defaults to 1) or null */
static int params_get(JNIEnv *env, jintArray params, slr_umi_assign_params *P)
{
    if (!params) return 0;
    const jsize len = (*env)->GetArrayLength(env, params);
    if (len < 5) return -1;
    jint *v = (*env)->GetIntArrayElements(env, params, NULL);
    if (!v) return -1;
    P->ed_complete = v[0]; P->ed_single = v[1]; P->single_threshold = v[2]; P->fold_depth = v[3]; P->max_hier = v[4];
    P->deep = len > 5 ? v[5] : 1;
    (*env)->ReleaseIntArrayElements(env, params, v, JNI_ABORT);
    return 1;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiDist(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                      jobject jobOffsets, jlong nJobs, jobject out, jobject outOffsets)
{
    (void)cls;
    umi_bufs U; int32_t *o; const int64_t *oo;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U) || !mat_check(env, &U, nJobs, out, outOffsets, 1, &o, &oo)) return SLR_E_INVALID;
    return slr_umi_dist(H(slr_ctx, ctx), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, o, oo);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiCluster(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                         jobject jobOffsets, jlong nJobs, jint ed, jobject member, jobject rank, jobject out,
                                                         jobject outOffsets, jobject rec)
{
    (void)cls;
    umi_bufs U; int32_t *o; const int64_t *oo;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U) || !mat_check(env, &U, nJobs, out, outOffsets, 0, &o, &oo)) return SLR_E_INVALID;
    int ok = 1;
    const uint8_t *mb = (const uint8_t *)buf(env, member, U.reads, 0, &ok);
    const int32_t *rk = (const int32_t *)buf(env, rank, U.reads * 4, 0, &ok);
    slr_umi_cluster_rec *r = (slr_umi_cluster_rec *)buf(env, rec, U.reads * (int64_t)sizeof(slr_umi_cluster_rec), 1, &ok);
    if (!ok) return SLR_E_INVALID;
    return slr_umi_cluster(H(slr_ctx, ctx), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, ed, mb, rk, o, oo, r);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiAssign(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                        jobject jobOffsets, jlong nJobs, jintArray params, jobject jobQv01, jobject out,
                                                        jobject outOffsets, jobject rec)
{
    (void)cls;
    umi_bufs U; int32_t *o; const int64_t *oo;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U) || !mat_check(env, &U, nJobs, out, outOffsets, 0, &o, &oo)) return SLR_E_INVALID;
    int ok = 1;
    slr_umi_assign_params P;
    const int hp = params_get(env, params, &P);
    const uint8_t *qv = (const uint8_t *)buf(env, jobQv01, (int64_t)nJobs, 0, &ok);
    slr_umi_assign_rec *r = (slr_umi_assign_rec *)buf(env, rec, U.reads * (int64_t)sizeof(slr_umi_assign_rec), 1, &ok);
    if (!ok || hp < 0) return SLR_E_INVALID;
    return slr_umi_assign(H(slr_ctx, ctx), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, hp ? &P : NULL, qv, o, oo, r);
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_umiSessionCreate(JNIEnv *env, jclass cls, jlong ctx, jobject umis, jint stride, jint umiLen,
                                                                jobject jobOffsets, jlong nJobs)
{
    (void)cls;
    umi_bufs U;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U)) return 0;
    slr_umi_session *s = NULL;
    const int rc = slr_umi_session_create(H(slr_ctx, ctx), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, &s);
    return rc == SLR_OK ? (jlong)(size_t)s : 0;
}

/* the session knows its own sizes: the record buffers are checked against slr_umi_session_reads / _jobs */
JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiSessionCluster(JNIEnv *env, jclass cls, jlong session, jint ed, jobject member, jobject rank,
                                                                jobject rec)
{
    (void)cls;
    if (!session) return SLR_E_INVALID;
    const int64_t m = slr_umi_session_reads(H(const slr_umi_session, session));
    int ok = 1;
    const uint8_t *mb = (const uint8_t *)buf(env, member, m, 0, &ok);
    const int32_t *rk = (const int32_t *)buf(env, rank, m * 4, 0, &ok);
    slr_umi_cluster_rec *r = (slr_umi_cluster_rec *)buf(env, rec, m * (int64_t)sizeof(slr_umi_cluster_rec), 1, &ok);
    if (!ok) return SLR_E_INVALID;
    return slr_umi_session_cluster(H(slr_umi_session, session), ed, mb, rk, r);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiSessionAssign(JNIEnv *env, jclass cls, jlong session, jintArray params, jobject jobQv01,
                                                               jobject rec)
{
    (void)cls;
    if (!session) return SLR_E_INVALID;
    const int64_t m = slr_umi_session_reads(H(const slr_umi_session, session));
    int ok = 1;
    slr_umi_assign_params P;
    const int hp = params_get(env, params, &P);
    const uint8_t *qv = (const uint8_t *)buf(env, jobQv01, slr_umi_session_jobs(H(const slr_umi_session, session)), 0, &ok);
    slr_umi_assign_rec *r = (slr_umi_assign_rec *)buf(env, rec, m * (int64_t)sizeof(slr_umi_assign_rec), 1, &ok);
    if (!ok || hp < 0) return SLR_E_INVALID;
    return slr_umi_session_assign(H(slr_umi_session, session), hp ? &P : NULL, qv, r);
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_umiSessionCells(JNIEnv *env, jclass cls, jlong session)
{
    (void)env; (void)cls;
    return (jlong)slr_umi_session_cells(H(const slr_umi_session, session));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_umiSessionMatrices(JNIEnv *env, jclass cls, jlong session, jobject out, jlong nCells)
{
    (void)cls;
    int ok = nCells >= 0;
    int32_t *o = (int32_t *)buf(env, out, (int64_t)nCells * 4, 1, &ok);
    if (!ok) return SLR_E_INVALID;
    return slr_umi_session_matrices(H(slr_umi_session, session), o, (int64_t)nCells);
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_umiSessionDestroy(JNIEnv *env, jclass cls, jlong session)
{
    (void)env; (void)cls;
    slr_umi_session_destroy(H(slr_umi_session, session));
}

/* ---- several GPUs behind this one JVM ------------------------------------------------------------------------------------------- */
JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_multiCreate(JNIEnv *env, jclass cls, jint nDevices, jintArray deviceIds, jint nStreams)
{
    (void)cls;
    jint *ids = NULL;
    if (deviceIds) {
        if ((*env)->GetArrayLength(env, deviceIds) < nDevices || nDevices <= 0) return 0;
        ids = (*env)->GetIntArrayElements(env, deviceIds, NULL);
        if (!ids) return 0;
    }
    slr_multi *m = NULL;
    const int rc = slr_multi_create(nDevices, (const int *)ids, nStreams, &m);
    if (ids) (*env)->ReleaseIntArrayElements(env, deviceIds, ids, JNI_ABORT);
    return rc == SLR_OK ? (jlong)(size_t)m : 0;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_multiDestroy(JNIEnv *env, jclass cls, jlong multi)
{
    (void)env; (void)cls;
    slr_multi_destroy(H(slr_multi, multi));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiDevices(JNIEnv *env, jclass cls, jlong multi)
{
    (void)env; (void)cls;
    return slr_multi_n_devices(H(const slr_multi, multi));
}

JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_multiBcTableCreate(JNIEnv *env, jclass cls, jlong multi, jlongArray barcodes, jintArray rank)
{
    (void)cls;
    return table_create(env, multi, 1, barcodes, rank);
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_multiBcTableDestroy(JNIEnv *env, jclass cls, jlong table)
{
    (void)env; (void)cls;
    slr_multi_bc_table_destroy(H(slr_multi_table, table));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiBcAssign(JNIEnv *env, jclass cls, jlong multi, jlong table, jint edMax, jint plusMinus,
                                                            jboolean threePrime, jobject slices, jint stride, jint sliceLen, jobject lens,
                                                            jobject anchor, jlong n, jobject out)
{
    (void)cls;
    bc_bufs B;
    if (!bc_check(env, slices, stride, lens, anchor, n, out, &B)) return SLR_E_INVALID;
    return slr_multi_bc_assign(H(slr_multi, multi), H(const slr_multi_table, table), edMax, plusMinus, threePrime ? 1 : 0, B.slices, stride, sliceLen,
                               B.lens, B.anchor, (int64_t)n, B.out);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiBcExact(JNIEnv *env, jclass cls, jlong multi, jlong table, jboolean threePrime, jobject slices,
                                                           jint stride, jint sliceLen, jobject lens, jobject anchor, jlong n, jobject out)
{
    (void)cls;
    bc_bufs B;
    if (!bc_check(env, slices, stride, lens, anchor, n, out, &B)) return SLR_E_INVALID;
    return slr_multi_bc_exact(H(slr_multi, multi), H(const slr_multi_table, table), threePrime ? 1 : 0, B.slices, stride, sliceLen, B.lens, B.anchor,
                              (int64_t)n, B.out);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiBcCountsRead(JNIEnv *env, jclass cls, jlong multi, jlong table, jlongArray countsOut)
{
    (void)cls;
    if (!multi || !table || !countsOut) return SLR_E_INVALID;
    int64_t *dc = NULL, ne = 0;
    slr_bc_table *t0 = slr_multi_bc_table_replica(H(slr_multi_table, table), 0);
    if (!t0 || slr_bc_counts_device(t0, &dc, &ne) != SLR_OK) return SLR_E_INVALID;
    if ((int64_t)(*env)->GetArrayLength(env, countsOut) < ne) return SLR_E_INVALID;
    jlong *c = (*env)->GetLongArrayElements(env, countsOut, NULL);
    if (!c) return SLR_E_NOMEM;
    const int rc = slr_multi_bc_counts_read(H(slr_multi, multi), H(slr_multi_table, table), (int64_t *)c);
    (*env)->ReleaseLongArrayElements(env, countsOut, c, 0);
    return rc;
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiBcCountsReset(JNIEnv *env, jclass cls, jlong multi, jlong table)
{
    (void)env; (void)cls;
    return slr_multi_bc_counts_reset(H(slr_multi, multi), H(slr_multi_table, table));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiUmiDist(JNIEnv *env, jclass cls, jlong multi, jobject umis, jint stride, jint umiLen,
                                                           jobject jobOffsets, jlong nJobs, jobject out, jobject outOffsets)
{
    (void)cls;
    umi_bufs U; int32_t *o; const int64_t *oo;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U) || !mat_check(env, &U, nJobs, out, outOffsets, 1, &o, &oo)) return SLR_E_INVALID;
    return slr_multi_umi_dist(H(slr_multi, multi), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, o, oo);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiUmiCluster(JNIEnv *env, jclass cls, jlong multi, jobject umis, jint stride, jint umiLen,
                                                              jobject jobOffsets, jlong nJobs, jint ed, jobject member, jobject rank, jobject rec)
{
    (void)cls;
    umi_bufs U;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U)) return SLR_E_INVALID;
    int ok = 1;
    const uint8_t *mb = (const uint8_t *)buf(env, member, U.reads, 0, &ok);
    const int32_t *rk = (const int32_t *)buf(env, rank, U.reads * 4, 0, &ok);
    slr_umi_cluster_rec *r = (slr_umi_cluster_rec *)buf(env, rec, U.reads * (int64_t)sizeof(slr_umi_cluster_rec), 1, &ok);
    if (!ok) return SLR_E_INVALID;
    return slr_multi_umi_cluster(H(slr_multi, multi), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, ed, mb, rk, r);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_multiUmiAssign(JNIEnv *env, jclass cls, jlong multi, jobject umis, jint stride, jint umiLen,
                                                             jobject jobOffsets, jlong nJobs, jintArray params, jobject jobQv01, jobject rec)
{
    (void)cls;
    umi_bufs U;
    if (!umi_check(env, umis, stride, jobOffsets, nJobs, &U)) return SLR_E_INVALID;
    int ok = 1;
    slr_umi_assign_params P;
    const int hp = params_get(env, params, &P);
    const uint8_t *qv = (const uint8_t *)buf(env, jobQv01, (int64_t)nJobs, 0, &ok);
    slr_umi_assign_rec *r = (slr_umi_assign_rec *)buf(env, rec, U.reads * (int64_t)sizeof(slr_umi_assign_rec), 1, &ok);
    if (!ok || hp < 0) return SLR_E_INVALID;
    return slr_multi_umi_assign(H(slr_multi, multi), U.umis, stride, umiLen, U.joff, (int64_t)nJobs, hp ? &P : NULL, qv, r);
}

/* ---- Illumina-guided search ------------------------------------------------------------------------------------------------------ */
JNIEXPORT jlong JNICALL Java_com_rw_gpu_Native_guidedSetsCreate(JNIEnv *env, jclass cls, jlong ctx, jlongArray groupKeys, jlongArray groupOffsets,
                                                                jlongArray allKeys, jint allEd, jlongArray emptyKeys, jint emptyEd,
                                                                jboolean bcFlavour, jint seqLen)
{
    (void)cls;
    if (!ctx || !groupOffsets) return 0;
    const jsize no = (*env)->GetArrayLength(env, groupOffsets);
    if (no < 1) return 0;
    const jsize ng = no - 1;
    jlong *go = (*env)->GetLongArrayElements(env, groupOffsets, NULL);
    if (!go) return 0;
    /* the key array must cover the last offset */
    const jsize nk = groupKeys ? (*env)->GetArrayLength(env, groupKeys) : 0;
    jlong res = 0;
    if (go[ng] <= (jlong)nk && go[0] >= 0) {
        jlong *gk = groupKeys ? (*env)->GetLongArrayElements(env, groupKeys, NULL) : NULL;
        jlong *ak = allKeys ? (*env)->GetLongArrayElements(env, allKeys, NULL) : NULL;
        jlong *ek = emptyKeys ? (*env)->GetLongArrayElements(env, emptyKeys, NULL) : NULL;
        if ((gk || !groupKeys) && (ak || !allKeys) && (ek || !emptyKeys)) {
            slr_guided_sets *s = NULL;
            const int rc = slr_guided_sets_create(H(slr_ctx, ctx), (const uint64_t *)gk, (const int64_t *)go, (int64_t)ng, (const uint64_t *)ak,
                                                  ak ? (int64_t)(*env)->GetArrayLength(env, allKeys) : 0, allEd, (const uint64_t *)ek,
                                                  ek ? (int64_t)(*env)->GetArrayLength(env, emptyKeys) : 0, emptyEd, bcFlavour ? 1 : 0, seqLen, &s);
            if (rc == SLR_OK) res = (jlong)(size_t)s;
        }
        if (gk) (*env)->ReleaseLongArrayElements(env, groupKeys, gk, JNI_ABORT);
        if (ak) (*env)->ReleaseLongArrayElements(env, allKeys, ak, JNI_ABORT);
        if (ek) (*env)->ReleaseLongArrayElements(env, emptyKeys, ek, JNI_ABORT);
    }
    (*env)->ReleaseLongArrayElements(env, groupOffsets, go, JNI_ABORT);
    return res;
}

JNIEXPORT void JNICALL Java_com_rw_gpu_Native_guidedSetsDestroy(JNIEnv *env, jclass cls, jlong sets)
{
    (void)env; (void)cls;
    slr_guided_sets_destroy(H(slr_guided_sets, sets));
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_guidedMatch(JNIEnv *env, jclass cls, jlong ctx, jlong sets, jint plusMinus, jint postLen, jint bailout,
                                                          jobject slices, jint stride, jint sliceLen, jobject anchor, jobject groupId, jobject ed,
                                                          jlong n, jobject out, jobject rawOut, jint rawCap)
{
    (void)cls;
    int ok = n >= 0 && stride > 0 && rawCap >= 0;
    const uint8_t *sl = (const uint8_t *)buf(env, slices, (int64_t)n * stride, 1, &ok);
    const int32_t *an = (const int32_t *)buf(env, anchor, (int64_t)n * 4, 1, &ok);
    const int32_t *gi = (const int32_t *)buf(env, groupId, (int64_t)n * 4, 1, &ok);
    const int32_t *e = (const int32_t *)buf(env, ed, (int64_t)n * 4, 1, &ok);
    slr_guided_result *o = (slr_guided_result *)buf(env, out, (int64_t)n * (int64_t)sizeof(slr_guided_result), 1, &ok);
    slr_guided_hit *ro = (slr_guided_hit *)buf(env, rawOut, (int64_t)n * rawCap * (int64_t)sizeof(slr_guided_hit), 0, &ok);
    if (!ok) return SLR_E_INVALID;
    return slr_guided_match(H(slr_ctx, ctx), H(const slr_guided_sets, sets), plusMinus, postLen, bailout, sl, stride, sliceLen, an, gi, e, (int64_t)n, o,
                            ro, rawCap);
}

/* scores7: NeedlemanScores {leading_gap_1, leading_gap_2, trailing_gap_1, trailing_gap_2, indel, mismatch, match} or null = the reference's defaults */
JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_guidedMismatchDiff(JNIEnv *env, jclass cls, jobject records, jlong n, jobject slices, jint stride,
                                                                 jint sliceLen, jobject anchor, jint seqLen, jintArray scores7, jobject diffOut)
{
    (void)cls;
    int ok = n >= 0 && stride > 0;
    const slr_guided_result *r = (const slr_guided_result *)buf(env, records, (int64_t)n * (int64_t)sizeof(slr_guided_result), 1, &ok);
    const uint8_t *sl = (const uint8_t *)buf(env, slices, (int64_t)n * stride, 1, &ok);
    const int32_t *an = (const int32_t *)buf(env, anchor, (int64_t)n * 4, 1, &ok);
    int32_t *d = (int32_t *)buf(env, diffOut, (int64_t)n * 4, 1, &ok);
    if (!ok) return SLR_E_INVALID;
    slr_needleman_scores sc;
    const slr_needleman_scores *psc = NULL;
    if (scores7) {
        if ((*env)->GetArrayLength(env, scores7) != 7) return SLR_E_INVALID;
        jint *v = (*env)->GetIntArrayElements(env, scores7, NULL);
        if (!v) return SLR_E_INVALID;
        sc.leading_gap_1 = v[0]; sc.leading_gap_2 = v[1]; sc.trailing_gap_1 = v[2]; sc.trailing_gap_2 = v[3]; sc.indel = v[4]; sc.mismatch = v[5]; sc.match = v[6];
        (*env)->ReleaseIntArrayElements(env, scores7, v, JNI_ABORT);
        psc = &sc;
    }
    return slr_guided_mismatch_diff(r, (int64_t)n, sl, stride, sliceLen, an, seqLen, psc, d);
}

JNIEXPORT jint JNICALL Java_com_rw_gpu_Native_dynMaxEd(JNIEnv *env, jclass cls, jlongArray maxCandidates, jint count, jint plusMinus, jint cap)
{
    (void)cls;
    if (!maxCandidates) return -1;
    const jsize n = (*env)->GetArrayLength(env, maxCandidates);
    jlong *m = (*env)->GetLongArrayElements(env, maxCandidates, NULL);
    if (!m) return -1;
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
