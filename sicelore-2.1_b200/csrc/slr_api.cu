// slr_api.cu — the C ABI of libsicelore_gpu.so (include/sicelore_gpu.h): context, device tables, batching.
// Host-side runtime only; the kernels are in bc_assign.cu / umi_dist.cu.  No CPU fallback anywhere: without a
// CUDA device every entry point fails with SLR_E_NODEVICE.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <memory>
#include <vector>
#include "slr_kernels.h"
#include "slr_table_build.h"
#include "umi_core.cuh"
#include "bc_core.cuh"
#include "guided_build.h"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess) return fail(SLR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// one kernel launch that needs a work-counter cell: `work` names the cell inside expr
#define CUDA_TRY_WORK(ctx_, stream_, expr)                                                                      \
    do {                                                                                                        \
        unsigned cell_;                                                                                         \
        unsigned long long *work = (ctx_)->work_acquire((stream_), cell_);                                      \
        cudaError_t e_ = (expr);                                                                                \
        (ctx_)->work_release((stream_), cell_);                                                                 \
        if (e_ != cudaSuccess) return fail(SLR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// reads per pipelined chunk of the host-pointer path (SLR_BC_CHUNK overrides, for tuning)
static long long bc_chunk()
{
    static const long long v = [] {
        const char *e = getenv("SLR_BC_CHUNK");
        const long long x = e ? atoll(e) : 0;
        return x >= 1024 ? x : (1LL << 18);         // 256 K reads: the UMI kernels of a concurrent caller interleave sooner (59.6 -> 55.7 ms per 10 M-read step)
    }();
    return v;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return SLR_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(SLR_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return SLR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// one slot = what one host thread needs to have a batch in flight: two streams + ping-pong device buffers
struct Slot {
    std::mutex mtx;
    cudaStream_t stream[2] = {nullptr, nullptr};
    DevBuf slices[2], anchor[2], lens[2], out[2];
    DevBuf umi[2], joff[2], ooff[2], uout[2], uscr[2];
    DevBuf jraw[2], jtmp[2];                   // the caller's job offsets of the range in flight + scan scratch (rebased on the device)
    DevBuf ucl[2];                             // slr_umi_cluster: counts | records | rank | member of the range in flight
    DevBuf uas[2];                             // slr_umi_assign: records | qv flags | job lists of the range in flight
    cudaEvent_t uscr_free = nullptr;           // recorded after the last launch that uses uscr[0] on a caller's stream (slr_umi_dist_dev)
    DevBuf gsl, ganc, ggid, ged, gout, graw, gvis;   // Illumina-guided search: staging buffers + the per-warp visited tables
    cudaEvent_t gvis_free = nullptr;           // recorded after the last guided launch that uses gvis
};

// Every error exit of a pipelined host-pointer call must leave nothing in flight: the D2H copy of the previous chunk may still be
// writing into the caller's (pinned) result buffer on the other ping-pong stream, and the caller is free to release it once we return.
struct SlotDrain {
    Slot *s;
    bool armed = true;
    explicit SlotDrain(Slot *slot) : s(slot) {}
    ~SlotDrain()
    {
        if (!armed) return;
        cudaStreamSynchronize(s->stream[0]);
        cudaStreamSynchronize(s->stream[1]);
    }
};

}  // namespace

constexpr unsigned WORK_COUNTERS = 1024;       // ring of 16-byte work-counter cells, one per kernel launch; a cell is handed out again only
                                               // behind the event of its previous user, whatever stream that launch ran on

struct slr_ctx {
    unsigned long long *d_work = nullptr;      // WORK_COUNTERS cells of two counters (umi_cluster uses both)
    cudaEvent_t work_done[WORK_COUNTERS] = {};
    std::mutex work_mtx;
    unsigned next_work = 0;
    // Cell for one launch on `stream`: the stream first waits for the launch that used the cell last (a no-op unless more than
    // WORK_COUNTERS launches are queued), the caller enqueues its reset + kernel, then calls work_release.
    unsigned long long *work_acquire(cudaStream_t stream, unsigned &cell)
    {
        std::lock_guard<std::mutex> lk(work_mtx);
        cell = next_work++ % WORK_COUNTERS;
        cudaStreamWaitEvent(stream, work_done[cell], 0);
        return d_work + 2 * (size_t)cell;
    }
    void work_release(cudaStream_t stream, unsigned cell)
    {
        std::lock_guard<std::mutex> lk(work_mtx);
        cudaEventRecord(work_done[cell], stream);
    }
    cudaStream_t aux = nullptr;                // counter read-back
    int device = 0;
    int n_slots = 1;
    std::vector<Slot *> slots;
    std::atomic<unsigned> next{0};
    std::mutex pool_mtx;                       // device buffers of destroyed UMI sessions, reused by the next one (cudaMalloc / cudaFree of
    std::vector<struct slr_umi_session *> session_pool;   // several GB cost more than the kernels)
};

static void free_session_pool(slr_ctx *c);     // defined with slr_umi_session below

struct slr_bc_table {
    slr_ctx *ctx = nullptr;
    SlrTableDev dev;
    void *d_buckets = nullptr;                 // 4 tables, contiguous
    void *d_stash_b = nullptr, *d_stash_s = nullptr;
    void *d_ix_keys = nullptr, *d_ix_vals = nullptr, *d_rank = nullptr, *d_counts = nullptr;
    long long n = 0, n_distinct = 0;
};

extern "C" {

const char *slr_last_error(void) { return g_err.c_str(); }
int slr_multi_fail(int code, const char *msg) { return fail(code, "%s", msg ? msg : ""); }     // slr_multi.cu reports through the same channel
int slr_abi_version(void) { return SLR_ABI_VERSION; }
int64_t slr_launch_count(void) { return g_launches.load(); }

int slr_ctx_create(int device, int n_streams, slr_ctx **out)
{
    if (!out) return fail(SLR_E_INVALID, "slr_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(SLR_E_NODEVICE, "no CUDA device available (%s); libsicelore_gpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return fail(SLR_E_INVALID, "device %d out of range (have %d)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0)   // the library holds sm_100a SASS only (no PTX): any other device would fail at the first launch
        return fail(SLR_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200, compute capability 10.0) only", device,
                    prop.major, prop.minor);
    slr_ctx *c = new slr_ctx();
    c->device = device;
    e = cudaMalloc((void **)&c->d_work, 2 * WORK_COUNTERS * sizeof(unsigned long long));
    if (e != cudaSuccess) { delete c; return fail(SLR_E_NOMEM, "cudaMalloc(work counters): %s", cudaGetErrorString(e)); }
    for (unsigned i = 0; i < WORK_COUNTERS && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->work_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking);
    if (e != cudaSuccess) { slr_ctx_destroy(c); return fail(SLR_E_CUDA, "slr_ctx_create: %s", cudaGetErrorString(e)); }
    c->n_slots = n_streams < 1 ? 1 : (n_streams > 64 ? 64 : n_streams);
    for (int i = 0; i < c->n_slots; i++) {
        Slot *s = new Slot();
        for (int k = 0; k < 2; k++) {
            e = cudaStreamCreateWithFlags(&s->stream[k], cudaStreamNonBlocking);
            if (e != cudaSuccess) { delete s; slr_ctx_destroy(c); return fail(SLR_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        }
        e = cudaEventCreateWithFlags(&s->uscr_free, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->gvis_free, cudaEventDisableTiming);
        if (e != cudaSuccess) { delete s; slr_ctx_destroy(c); return fail(SLR_E_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); }
        c->slots.push_back(s);
    }
    *out = c;
    return SLR_OK;
}

void slr_ctx_destroy(slr_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (Slot *s : c->slots) {
        for (int k = 0; k < 2; k++) {
            if (s->stream[k]) { cudaStreamSynchronize(s->stream[k]); cudaStreamDestroy(s->stream[k]); }
            s->slices[k].release(); s->anchor[k].release(); s->lens[k].release(); s->out[k].release();
        }
        for (int k = 0; k < 2; k++) { s->umi[k].release(); s->joff[k].release(); s->ooff[k].release(); s->uout[k].release(); s->uscr[k].release(); s->ucl[k].release(); s->uas[k].release(); s->jraw[k].release(); s->jtmp[k].release(); }
        if (s->uscr_free) cudaEventDestroy(s->uscr_free);
        if (s->gvis_free) cudaEventDestroy(s->gvis_free);
        s->gsl.release(); s->ganc.release(); s->ggid.release(); s->ged.release(); s->gout.release(); s->graw.release(); s->gvis.release();
        delete s;
    }
    free_session_pool(c);
    cudaDeviceSynchronize();                   // launches of the *_dev entry points on the caller's streams may still use the counter cells
    for (unsigned i = 0; i < WORK_COUNTERS; i++)
        if (c->work_done[i]) cudaEventDestroy(c->work_done[i]);
    if (c->aux) cudaStreamDestroy(c->aux);
    cudaFree(c->d_work);
    delete c;
}

int slr_ctx_device(const slr_ctx *c) { return c ? c->device : -1; }

// ---------------------------------------------------------------------------------------------------------
int slr_bc_table_create(slr_ctx *ctx, const uint64_t *barcodes2bit, const int32_t *rank, int64_t n, int bc_len, slr_bc_table **out)
{
    if (!ctx || !out || (!barcodes2bit && n > 0) || n < 0) return fail(SLR_E_INVALID, "slr_bc_table_create: bad argument");
    *out = nullptr;
    if (bc_len != 16) return fail(SLR_E_UNSUPPORTED, "cell_bc_length %d not supported (only 16)", bc_len);
    if (n > 0x7FFFFFFFLL) return fail(SLR_E_UNSUPPORTED, "more than 2^31 barcodes");
    CUDA_TRY(cudaSetDevice(ctx->device));
    SlrTableHost H;
    slr_build_table(barcodes2bit, rank, n, H);
    slr_bc_table *t = new slr_bc_table();
    t->ctx = ctx; t->n = n; t->n_distinct = H.n_distinct;
    memset(&t->dev, 0, sizeof(t->dev));
    const size_t tbytes = (size_t)32 << H.bbits;        // 32-byte buckets
#define TRY_OR_FREE(expr)                                                                                       \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess) { slr_bc_table_destroy(t); return fail(SLR_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } \
    } while (0)
    TRY_OR_FREE(cudaMalloc(&t->d_buckets, 4 * tbytes));
    TRY_OR_FREE(cudaMemcpy(t->d_buckets, H.slots.data(), 4 * tbytes, cudaMemcpyHostToDevice));
    t->dev.bk = reinterpret_cast<const uint4 *>(t->d_buckets);
    const size_t sn = H.st_bucket.size();
    t->dev.st_total = (int)sn;
    if (sn) {
        TRY_OR_FREE(cudaMalloc(&t->d_stash_b, sn * 4));
        TRY_OR_FREE(cudaMalloc(&t->d_stash_s, sn * 2));
        TRY_OR_FREE(cudaMemcpy(t->d_stash_b, H.st_bucket.data(), sn * 4, cudaMemcpyHostToDevice));
        TRY_OR_FREE(cudaMemcpy(t->d_stash_s, H.st_slot.data(), sn * 2, cudaMemcpyHostToDevice));
    }
    t->dev.st_bucket = (const uint32_t *)t->d_stash_b;
    t->dev.st_slot = (const uint16_t *)t->d_stash_s;
    t->dev.bbits = H.bbits;
    const size_t ixn = (size_t)H.ix_mask + 1;
    TRY_OR_FREE(cudaMalloc(&t->d_ix_keys, ixn * 4));
    TRY_OR_FREE(cudaMalloc(&t->d_ix_vals, ixn * 4));
    TRY_OR_FREE(cudaMemcpy(t->d_ix_keys, H.ix_keys.data(), ixn * 4, cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMemcpy(t->d_ix_vals, H.ix_vals.data(), ixn * 4, cudaMemcpyHostToDevice));
    t->dev.ix_keys = (const uint32_t *)t->d_ix_keys;
    t->dev.ix_vals = (const int32_t *)t->d_ix_vals;
    t->dev.ix_mask = H.ix_mask;
    if (rank && n > 0) {
        TRY_OR_FREE(cudaMalloc(&t->d_rank, (size_t)n * 4));
        TRY_OR_FREE(cudaMemcpy(t->d_rank, rank, (size_t)n * 4, cudaMemcpyHostToDevice));
    }
    t->dev.rank = (const int32_t *)t->d_rank;
    const size_t cbytes = (size_t)(n > 0 ? n : 1) * 3 * sizeof(unsigned long long);
    TRY_OR_FREE(cudaMalloc(&t->d_counts, cbytes));
    TRY_OR_FREE(cudaMemset(t->d_counts, 0, cbytes));
    t->dev.counts = (unsigned long long *)t->d_counts;
    t->dev.n = n;
#undef TRY_OR_FREE
    *out = t;
    return SLR_OK;
}

void slr_bc_table_destroy(slr_bc_table *t)
{
    if (!t) return;
    if (t->ctx) cudaSetDevice(t->ctx->device);
    cudaFree(t->d_buckets);
    cudaFree(t->d_stash_b); cudaFree(t->d_stash_s);
    cudaFree(t->d_ix_keys); cudaFree(t->d_ix_vals); cudaFree(t->d_rank); cudaFree(t->d_counts);
    delete t;
}

int64_t slr_bc_table_size(const slr_bc_table *t) { return t ? t->n_distinct : 0; }

// ---------------------------------------------------------------------------------------------------------
static int check_bc_args(const slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int stride, int slice_len, int64_t n)
{
    if (!ctx || !t) return fail(SLR_E_INVALID, "slr_bc_assign: ctx / table is NULL");
    if (t->ctx != ctx) return fail(SLR_E_INVALID, "slr_bc_assign: the table belongs to another context (device %d, this context: device %d)",
                                   t->ctx ? t->ctx->device : -1, ctx->device);
    if (ed_max < 0) return fail(SLR_E_INVALID, "bcEditDistance %d < 0", ed_max);
    if (ed_max > 2) return fail(SLR_E_UNSUPPORTED, "bcEditDistance %d not supported by the GPU path (0, 1 or 2)", ed_max);
    if (plusminus < 0 || 2 * plusminus + 1 > SLR_MAX_OFFSETS) return fail(SLR_E_UNSUPPORTED, "testPlusMinusPos %d not supported (0..4)", plusminus);
    if (slice_len < 1 || slice_len > 32 || stride < slice_len) return fail(SLR_E_INVALID, "slice_len %d / stride %d invalid (1 <= slice_len <= 32, slice_len <= stride)", slice_len, stride);
    if (n < 0) return fail(SLR_E_INVALID, "n < 0");
    return SLR_OK;
}

static int bc_assign_dev_impl(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime, int need_post,
                              const uint8_t *d_slices, int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor, int64_t n,
                              slr_bc_result *d_out, void *stream)
{
    int rc = check_bc_args(ctx, t, ed_max, plusminus, stride, slice_len, n);
    if (rc) return rc;
    if (n == 0) return SLR_OK;
    if (!d_slices || !d_anchor || !d_out) return fail(SLR_E_INVALID, "slr_bc_assign_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int64_t off = 0; off < n; off += (1LL << 30)) {           // grid.x limit: 2^31-1 blocks of 8 reads
        const int64_t m = (n - off) < (1LL << 30) ? (n - off) : (1LL << 30);
        CUDA_TRY_WORK(ctx, (cudaStream_t)stream,
                      slr_launch_bc_assign(t->dev, ed_max, plusminus, three_prime, need_post, d_slices + off * stride, stride, slice_len,
                                           d_lens ? d_lens + off : nullptr, d_anchor + off, m, d_out + off, work, (cudaStream_t)stream));
        g_launches++;
    }
    return SLR_OK;
}

int slr_bc_assign_dev(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime, const uint8_t *d_slices,
                      int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor, int64_t n, slr_bc_result *d_out,
                      void *stream)
{
    return bc_assign_dev_impl(ctx, t, ed_max, plusminus, three_prime, 1, d_slices, stride, slice_len, d_lens, d_anchor, n, d_out, stream);
}

int slr_bc_exact_dev(slr_ctx *ctx, const slr_bc_table *t, int three_prime, const uint8_t *d_slices, int stride, int slice_len,
                     const int32_t *d_lens, const int32_t *d_anchor, int64_t n, slr_bc_result *d_out, void *stream)
{
    return bc_assign_dev_impl(ctx, t, 0, 0, three_prime, 0, d_slices, stride, slice_len, d_lens, d_anchor, n, d_out, stream);
}

static int bc_assign_host_impl(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime, int need_post,
                               const uint8_t *slices, int stride, int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n,
                               slr_bc_result *out)
{
    int rc = check_bc_args(ctx, t, ed_max, plusminus, stride, slice_len, n);
    if (rc) return rc;
    if (n == 0) return SLR_OK;
    if (!slices || !anchor || !out) return fail(SLR_E_INVALID, "slr_bc_assign: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *s = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(s->mtx);
    SlotDrain drain(s);                                                // on any error exit: nothing left in flight on the caller's buffers
    // chunked ping-pong pipeline: H2D of chunk c+1 (stream B) overlaps the kernel of chunk c (stream A)
    int c = 0;
    const int64_t BC_CHUNK = bc_chunk();
    for (int64_t off = 0; off < n; off += BC_CHUNK, c++) {
        const int64_t m = (n - off) < BC_CHUNK ? (n - off) : BC_CHUNK;
        const int b = c & 1;
        cudaStream_t st = s->stream[b];
        CUDA_TRY(cudaStreamSynchronize(st));                       // buffers of this parity are free again
        if ((rc = s->slices[b].reserve((size_t)m * stride))) return rc;
        if ((rc = s->anchor[b].reserve((size_t)m * 4))) return rc;
        if ((rc = s->out[b].reserve((size_t)m * sizeof(slr_bc_result)))) return rc;
        if (lens && (rc = s->lens[b].reserve((size_t)m * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->slices[b].p, slices + off * stride, (size_t)m * stride, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s->anchor[b].p, anchor + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        if (lens) CUDA_TRY(cudaMemcpyAsync(s->lens[b].p, lens + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY_WORK(ctx, st, slr_launch_bc_assign(t->dev, ed_max, plusminus, three_prime, need_post, (const uint8_t *)s->slices[b].p, stride,
                                                    slice_len, lens ? (const int32_t *)s->lens[b].p : nullptr, (const int32_t *)s->anchor[b].p, m,
                                                    (slr_bc_result *)s->out[b].p, work, st));
        g_launches++;
        CUDA_TRY(cudaMemcpyAsync(out + off, s->out[b].p, (size_t)m * sizeof(slr_bc_result), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(s->stream[1]));
    return SLR_OK;
}

int slr_bc_assign(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime, const uint8_t *slices, int stride,
                  int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out)
{
    return bc_assign_host_impl(ctx, t, ed_max, plusminus, three_prime, 1, slices, stride, slice_len, lens, anchor, n, out);
}

int slr_bc_exact(slr_ctx *ctx, const slr_bc_table *t, int three_prime, const uint8_t *slices, int stride, int slice_len,
                 const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out)
{
    return bc_assign_host_impl(ctx, t, 0, 0, three_prime, 0, slices, stride, slice_len, lens, anchor, n, out);
}

int slr_bc_counts_read(slr_ctx *ctx, const slr_bc_table *t, int64_t *counts_out)
{
    if (!ctx || !t || !counts_out) return fail(SLR_E_INVALID, "slr_bc_counts_read: NULL argument");
    if (t->ctx != ctx) return fail(SLR_E_INVALID, "slr_bc_counts_read: the table belongs to another context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    // The copy is ordered behind everything the host-pointer calls have enqueued on the context's slot streams so far (an event per
    // stream, no device-wide synchronisation: callers that are still submitting are not held up).  Launches of the *_dev entry points
    // run on the caller's streams: the caller synchronises those before reading the counters.
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaSuccess;
    for (Slot *s : ctx->slots)
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            e = cudaEventRecord(ev, s->stream[k]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->aux, ev, 0);
        }
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts_out, t->d_counts, (size_t)t->n * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->aux);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->aux);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(SLR_E_CUDA, "slr_bc_counts_read: %s", cudaGetErrorString(e));
    return SLR_OK;
}

int slr_bc_counts_reset(slr_ctx *ctx, slr_bc_table *t)
{
    if (!ctx || !t) return fail(SLR_E_INVALID, "slr_bc_counts_reset: NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());                                 // every launch that still adds to the counters has finished
    CUDA_TRY(cudaMemsetAsync(t->d_counts, 0, (size_t)(t->n > 0 ? t->n : 1) * 3 * sizeof(int64_t), ctx->aux));
    CUDA_TRY(cudaStreamSynchronize(ctx->aux));                         // (a plain cudaMemset is not ordered against the non-blocking slot streams)
    return SLR_OK;
}

int slr_bc_counts_device(const slr_bc_table *t, int64_t **d_counts, int64_t *n_elems)
{
    if (!t || !d_counts || !n_elems) return fail(SLR_E_INVALID, "slr_bc_counts_device: NULL argument");
    *d_counts = (int64_t *)t->d_counts;
    *n_elems = t->n * 3;
    return SLR_OK;
}

// ---------------------------------------------------------------------------------------------------------
int slr_bc_collide_dev(slr_ctx *ctx, const slr_bc_table *t, int ed_max, const uint64_t *d_barcodes, int64_t n, slr_collide_result *d_out,
                       void *stream)
{
    if (!ctx || !t) return fail(SLR_E_INVALID, "slr_bc_collide: ctx / table is NULL");
    if (t->ctx != ctx) return fail(SLR_E_INVALID, "slr_bc_collide: the table belongs to another context");
    if (ed_max < 0 || n < 0) return fail(SLR_E_INVALID, "slr_bc_collide: mergeBCsED %d / n %lld invalid", ed_max, (long long)n);
    if (ed_max > 2) return fail(SLR_E_UNSUPPORTED, "mergeBCsED %d not supported by the GPU path (0, 1 or 2)", ed_max);
    if (n == 0) return SLR_OK;
    if (!d_barcodes || !d_out) return fail(SLR_E_INVALID, "slr_bc_collide_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(slr_launch_bc_collide(t->dev, ed_max, (const unsigned long long *)d_barcodes, n, d_out, (cudaStream_t)stream));
    g_launches++;
    return SLR_OK;
}

int slr_bc_collide(slr_ctx *ctx, const slr_bc_table *t, int ed_max, const uint64_t *barcodes, int64_t n, slr_collide_result *out)
{
    if (!ctx || !t) return fail(SLR_E_INVALID, "slr_bc_collide: ctx / table is NULL");
    if (t->ctx != ctx) return fail(SLR_E_INVALID, "slr_bc_collide: the table belongs to another context");
    if (n > 0 && (!barcodes || !out)) return fail(SLR_E_INVALID, "slr_bc_collide: NULL buffer");
    if (ed_max < 0 || n < 0) return fail(SLR_E_INVALID, "slr_bc_collide: mergeBCsED %d / n %lld invalid", ed_max, (long long)n);
    if (ed_max > 2) return fail(SLR_E_UNSUPPORTED, "mergeBCsED %d not supported by the GPU path (0, 1 or 2)", ed_max);
    if (n == 0) return SLR_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *s = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(s->mtx);
    cudaStream_t st = s->stream[0];
    CUDA_TRY(cudaStreamSynchronize(st));
    int rc;
    if ((rc = s->slices[0].reserve((size_t)n * 8))) return rc;             // the slot's staging buffers double as query / result space
    if ((rc = s->out[0].reserve((size_t)n * sizeof(slr_collide_result)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->slices[0].p, barcodes, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(slr_launch_bc_collide(t->dev, ed_max, (const unsigned long long *)s->slices[0].p, n, (slr_collide_result *)s->out[0].p, st));
    g_launches++;
    CUDA_TRY(cudaMemcpyAsync(out, s->out[0].p, (size_t)n * sizeof(slr_collide_result), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SLR_OK;
}

// ---------------------------------------------------------------------------------------------------------
static int check_umi_args(const slr_ctx *ctx, int stride, int umi_len, int64_t n_jobs)
{
    if (!ctx) return fail(SLR_E_INVALID, "slr_umi_dist: ctx is NULL");
    if (umi_len < 1 || umi_len > SLR_UMI_MAX_LEN) return fail(SLR_E_UNSUPPORTED, "umi_length %d not supported (1..%d)", umi_len, SLR_UMI_MAX_LEN);
    if (stride < umi_len + 2) return fail(SLR_E_INVALID, "stride %d < umi_len + 2", stride);
    if (n_jobs < 0) return fail(SLR_E_INVALID, "n_jobs < 0");
    return SLR_OK;
}

int slr_umi_dist_dev(slr_ctx *ctx, const uint8_t *d_umis, int stride, int umi_len, const int64_t *d_job_offsets, int64_t n_jobs,
                     int64_t n_reads, int32_t *d_out, const int64_t *d_out_offsets, int64_t n_out, void *stream)
{
    int rc = check_umi_args(ctx, stride, umi_len, n_jobs);
    if (rc) return rc;
    (void)n_out;
    if (n_jobs == 0 || n_reads == 0) return SLR_OK;
    if (!d_umis || !d_job_offsets || !d_out || !d_out_offsets) return fail(SLR_E_INVALID, "slr_umi_dist_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    // the kernels need a per-read scratch area: take a slot's, ordered behind its previous user with an event
    Slot *s = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(s->mtx);
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->uscr_free, 0));
    if ((rc = s->uscr[0].reserve(slr_umi_scratch_bytes(n_reads)))) return rc;
    CUDA_TRY(slr_launch_umi_dist(d_umis, stride, umi_len, (const long long *)d_job_offsets, n_jobs, n_reads, d_out,
                                 (const long long *)d_out_offsets, s->uscr[0].p, (cudaStream_t)stream));
    CUDA_TRY(cudaEventRecord(s->uscr_free, (cudaStream_t)stream));
    g_launches += SLR_UMI_LAUNCHES;
    return SLR_OK;
}

}  // extern "C"

namespace {
struct ClusterArgs {                           // slr_umi_cluster rides on the range loop of slr_umi_dist
    int ed = 0;
    const uint8_t *member = nullptr;
    const int32_t *rank = nullptr;
    slr_umi_cluster_rec *rec = nullptr;
};
struct AssignArgs {                            // slr_umi_assign: the same ride
    slr_umi_assign_params P;
    const uint8_t *job_qv01 = nullptr;
    slr_umi_assign_rec *rec = nullptr;
};
const slr_umi_assign_params ASSIGN_DEFAULTS = {2, 1, 3000, 50, 100, 1};   // config.xml:270-278, UMIparameters.java:L118, UmiClustering.java:L240

int check_assign_params(const slr_umi_assign_params *p, slr_umi_assign_params &out)
{
    out = p ? *p : ASSIGN_DEFAULTS;
    if (out.ed_complete < 0 || out.ed_complete > 5 || out.ed_single < 0 || out.ed_single > 5)
        return fail(SLR_E_INVALID, "slr_umi_assign: clustering edit distances %d / %d outside 0..5", out.ed_complete, out.ed_single);
    if (out.fold_depth < 1) return fail(SLR_E_INVALID, "slr_umi_assign: foldDepthBelowMaxDiscardForClustering %d < 1", out.fold_depth);
    if (out.max_hier < 0 || out.max_hier > 100)
        return fail(SLR_E_UNSUPPORTED, "slr_umi_assign: max_hier %d (ClusterOneHierarchical takes jobs of at most 100 reads)", out.max_hier);
    if (out.single_threshold < 0) return fail(SLR_E_INVALID, "slr_umi_assign: negative single-link threshold");
    if (out.deep != 0 && out.deep != 1) return fail(SLR_E_INVALID, "slr_umi_assign: deep must be 0 or 1");
    return SLR_OK;
}
// working arrays of the jobs ClusterOne_MyClustering gets (umi_assign_deep.cu), in 32-bit words
long long deep_words_of(const int64_t *job_offsets, int64_t j0, int64_t j1, const slr_umi_assign_params &P)
{
    if (!P.deep) return 0;
    long long w = 0;
    for (int64_t k = j0; k < j1; k++) {
        const int64_t n = job_offsets[k + 1] - job_offsets[k];
        if (n > P.max_hier) w += slr_umi_assign_deep_words(n);
    }
    return w;
}
}  // namespace

static int umi_dist_ranges(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                           int32_t *out, const int64_t *out_offsets, const ClusterArgs *cl, const AssignArgs *as = nullptr)
{
    int rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *s = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(s->mtx);
    SlotDrain drain(s);                                                    // on any error exit: nothing left in flight on the caller's buffers
    // job ranges of at most ~2^23 output cells (a single larger job still goes in one launch), ping-pong on the slot's two
    // streams: H2D of range c+1 and D2H of range c-1 overlap the kernels of range c.  The host only walks the job sizes to cut the
    // ranges; the offsets relative to the range and the matrix offsets (prefix sums of n^2) are made on the device.
    const int64_t CELL_LIMIT = 1LL << 23;
    std::vector<long long> ooff[2];                                        // only for a caller whose matrices are not packed back to back
    CUDA_TRY(cudaStreamWaitEvent(s->stream[0], s->uscr_free, 0));         // uscr[0] may still serve a slr_umi_dist_dev launch
    int64_t j = 0;
    for (int c = 0; j < n_jobs; c++) {
        const int b = c & 1;
        cudaStream_t st = s->stream[b];
        CUDA_TRY(cudaStreamSynchronize(st));                               // buffers (and host vectors) of this parity are free again
        int64_t j1 = j, cells = 0;
        bool packed = true;                                                // out_offsets of the range = running sum of n^2
        const int64_t r0 = job_offsets[j];
        while (j1 < n_jobs) {
            const int64_t nj = job_offsets[j1 + 1] - job_offsets[j1];
            if (nj < 0) return fail(SLR_E_INVALID, "job_offsets not monotone at %lld", (long long)j1);
            if (j1 > j && cells + nj * nj > CELL_LIMIT) break;
            if (out) packed &= out_offsets[j1] - out_offsets[j] == cells;
            cells += nj * nj;
            j1++;
        }
        const int64_t nr = job_offsets[j1] - r0, nj_range = j1 - j;
        if (nr > 0) {
            const size_t off_bytes = (size_t)(nj_range + 1) * 8;
            if ((rc = s->umi[b].reserve((size_t)nr * stride))) return rc;
            if ((rc = s->jraw[b].reserve(off_bytes))) return rc;
            if ((rc = s->jtmp[b].reserve(slr_umi_rebase_tmp_bytes(nj_range)))) return rc;
            if ((rc = s->joff[b].reserve(off_bytes))) return rc;
            if ((rc = s->ooff[b].reserve(off_bytes))) return rc;
            if ((rc = s->uout[b].reserve((size_t)cells * 4))) return rc;
            if ((rc = s->uscr[b].reserve(slr_umi_scratch_bytes(nr)))) return rc;
            CUDA_TRY(cudaMemcpyAsync(s->umi[b].p, umis + r0 * stride, (size_t)nr * stride, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(s->jraw[b].p, job_offsets + j, off_bytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(slr_launch_umi_rebase((const long long *)s->jraw[b].p, nj_range, r0, (long long *)s->joff[b].p, (long long *)s->ooff[b].p,
                                           s->jtmp[b].p, st));
            g_launches += SLR_UMI_REBASE_LAUNCHES;
            CUDA_TRY(slr_launch_umi_dist((const uint8_t *)s->umi[b].p, stride, umi_len, (const long long *)s->joff[b].p, nj_range, nr,
                                         (int32_t *)s->uout[b].p, (const long long *)s->ooff[b].p, s->uscr[b].p, st));
            g_launches += SLR_UMI_LAUNCHES;
            if (cl) {
                const size_t o_rec = ((size_t)nr * 4 + 15) & ~(size_t)15, o_rank = o_rec + (size_t)nr * 16, o_mem = o_rank + (size_t)nr * 4;
                if ((rc = s->ucl[b].reserve(o_mem + (size_t)nr))) return rc;
                char *base = (char *)s->ucl[b].p;
                if (cl->rank) CUDA_TRY(cudaMemcpyAsync(base + o_rank, cl->rank + r0, (size_t)nr * 4, cudaMemcpyHostToDevice, st));
                if (cl->member) CUDA_TRY(cudaMemcpyAsync(base + o_mem, cl->member + r0, (size_t)nr, cudaMemcpyHostToDevice, st));
                CUDA_TRY_WORK(ctx, st, slr_launch_umi_cluster((const int32_t *)s->uout[b].p, (const long long *)s->joff[b].p,
                                                              (const long long *)s->ooff[b].p, nj_range, nr, cl->ed,
                                                              cl->member ? (const uint8_t *)(base + o_mem) : nullptr,
                                                              cl->rank ? (const int32_t *)(base + o_rank) : nullptr,
                                                              slr_umi_scratch_rowjob(s->uscr[b].p, nr), (int32_t *)base,
                                                              (slr_umi_cluster_rec *)(base + o_rec), work, st));
                g_launches += SLR_UMI_CLUSTER_LAUNCHES;
                CUDA_TRY(cudaMemcpyAsync(cl->rec + r0, base + o_rec, (size_t)nr * 16, cudaMemcpyDeviceToHost, st));
            }
            if (as) {
                const size_t o_qv = (size_t)nr * sizeof(slr_umi_assign_rec), o_list = (o_qv + (size_t)nj_range + 15) & ~(size_t)15;
                const long long dw = deep_words_of(job_offsets, j, j1, as->P);
                const size_t scr_bytes = slr_umi_assign_scratch(nj_range, dw);
                if ((rc = s->uas[b].reserve(o_list + scr_bytes))) return rc;
                char *base = (char *)s->uas[b].p;
                if (as->job_qv01) CUDA_TRY(cudaMemcpyAsync(base + o_qv, as->job_qv01 + j, (size_t)nj_range, cudaMemcpyHostToDevice, st));
                CUDA_TRY(slr_launch_umi_assign((const int32_t *)s->uout[b].p, (const long long *)s->joff[b].p, (const long long *)s->ooff[b].p,
                                               nj_range, nr, as->P, as->job_qv01 ? (const uint8_t *)(base + o_qv) : nullptr,
                                               slr_umi_scratch_rowjob(s->uscr[b].p, nr), (slr_umi_assign_rec *)base, base + o_list, scr_bytes, st));
                g_launches += SLR_UMI_ASSIGN_LAUNCHES + (dw > 0 ? SLR_UMI_ASSIGN_DEEP_LAUNCHES : 0);
                CUDA_TRY(cudaMemcpyAsync(as->rec + r0, base, (size_t)nr * sizeof(slr_umi_assign_rec), cudaMemcpyDeviceToHost, st));
            }
            if (out && packed) {                                           // the usual layout: the range is one contiguous piece of `out`
                if (cells > 0)
                    CUDA_TRY(cudaMemcpyAsync(out + out_offsets[j], s->uout[b].p, (size_t)cells * 4, cudaMemcpyDeviceToHost, st));
            } else if (out) {                                              // jobs may sit anywhere in the caller's `out`: one copy per contiguous run
                ooff[b].clear();
                int64_t acc = 0;
                for (int64_t k = j; k < j1; k++) {
                    ooff[b].push_back(acc);
                    const int64_t nk = job_offsets[k + 1] - job_offsets[k];
                    acc += nk * nk;
                }
                int64_t a = j;
                while (a < j1) {
                    int64_t e = a;
                    while (e + 1 < j1) {
                        const int64_t ne = job_offsets[e + 1] - job_offsets[e];
                        if (out_offsets[e + 1] != out_offsets[e] + ne * ne) break;
                        e++;
                    }
                    const int64_t ne = job_offsets[e + 1] - job_offsets[e];
                    const int64_t ncell = ooff[b][e - j] + ne * ne - ooff[b][a - j];
                    if (ncell > 0)
                        CUDA_TRY(cudaMemcpyAsync(out + out_offsets[a], (int32_t *)s->uout[b].p + ooff[b][a - j], (size_t)ncell * 4,
                                                 cudaMemcpyDeviceToHost, st));
                    a = e + 1;
                }
            }
        }
        j = j1;
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(s->stream[1]));
    return SLR_OK;
}

extern "C" {

int slr_umi_dist(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int32_t *out,
                 const int64_t *out_offsets)
{
    int rc = check_umi_args(ctx, stride, umi_len, n_jobs);
    if (rc) return rc;
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !out || !out_offsets) return fail(SLR_E_INVALID, "slr_umi_dist: NULL buffer");
    return umi_dist_ranges(ctx, umis, stride, umi_len, job_offsets, n_jobs, out, out_offsets, nullptr);
}

int slr_umi_cluster(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int ed,
                    const uint8_t *member, const int32_t *rank, int32_t *out, const int64_t *out_offsets, slr_umi_cluster_rec *rec)
{
    int rc = check_umi_args(ctx, stride, umi_len, n_jobs);
    if (rc) return rc;
    if (ed < 0 || ed > 5) return fail(SLR_E_INVALID, "slr_umi_cluster: ed %d outside 0..5", ed);
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !rec || (out && !out_offsets)) return fail(SLR_E_INVALID, "slr_umi_cluster: NULL buffer");
    ClusterArgs cl;
    cl.ed = ed; cl.member = member; cl.rank = rank; cl.rec = rec;
    return umi_dist_ranges(ctx, umis, stride, umi_len, job_offsets, n_jobs, out, out_offsets, &cl);
}

int slr_umi_assign(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                   const slr_umi_assign_params *params, const uint8_t *job_qv01, int32_t *out, const int64_t *out_offsets,
                   slr_umi_assign_rec *rec)
{
    int rc = check_umi_args(ctx, stride, umi_len, n_jobs);
    if (rc) return rc;
    AssignArgs as;
    if ((rc = check_assign_params(params, as.P))) return rc;
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !rec || (out && !out_offsets)) return fail(SLR_E_INVALID, "slr_umi_assign: NULL buffer");
    as.job_qv01 = job_qv01; as.rec = rec;
    return umi_dist_ranges(ctx, umis, stride, umi_len, job_offsets, n_jobs, out, out_offsets, nullptr, &as);
}

int64_t slr_umi_assign_scratch_bytes(int64_t n_jobs) { return (int64_t)slr_umi_assign_scratch(n_jobs > 0 ? n_jobs : 0, 0); }
int64_t slr_umi_assign_deep_job_bytes(int64_t n_reads_of_job) { return n_reads_of_job > 0 ? 4 * (int64_t)slr_umi_assign_deep_words(n_reads_of_job) : 0; }

int slr_umi_assign_dev(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets, const int64_t *d_out_offsets, int64_t n_jobs,
                       int64_t n_reads, const slr_umi_assign_params *params, const uint8_t *d_job_qv01, void *d_scratch,
                       slr_umi_assign_rec *d_rec, void *stream)
{
    return slr_umi_assign_dev2(ctx, d_matrices, d_job_offsets, d_out_offsets, n_jobs, n_reads, params, d_job_qv01, d_scratch,
                               slr_umi_assign_scratch_bytes(n_jobs), d_rec, stream);
}

int slr_umi_assign_dev2(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets, const int64_t *d_out_offsets, int64_t n_jobs,
                        int64_t n_reads, const slr_umi_assign_params *params, const uint8_t *d_job_qv01, void *d_scratch, int64_t scratch_bytes,
                        slr_umi_assign_rec *d_rec, void *stream)
{
    if (!ctx) return fail(SLR_E_INVALID, "slr_umi_assign_dev: ctx is NULL");
    slr_umi_assign_params P;
    int rc = check_assign_params(params, P);
    if (rc) return rc;
    if (n_jobs < 0 || n_reads < 0) return fail(SLR_E_INVALID, "slr_umi_assign_dev: negative size");
    if (n_jobs == 0 || n_reads == 0) return SLR_OK;
    if (!d_matrices || !d_job_offsets || !d_out_offsets || !d_scratch || !d_rec) return fail(SLR_E_INVALID, "slr_umi_assign_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (scratch_bytes < slr_umi_assign_scratch_bytes(n_jobs))
        return fail(SLR_E_INVALID, "slr_umi_assign_dev: %lld bytes of scratch, %lld needed for the job lists alone", (long long)scratch_bytes,
                    (long long)slr_umi_assign_scratch_bytes(n_jobs));
    const bool deep = P.deep && scratch_bytes > slr_umi_assign_scratch_bytes(n_jobs);
    CUDA_TRY(slr_launch_umi_assign(d_matrices, (const long long *)d_job_offsets, (const long long *)d_out_offsets, n_jobs, n_reads, P, d_job_qv01,
                                   nullptr, d_rec, d_scratch, (size_t)scratch_bytes, (cudaStream_t)stream));
    g_launches += SLR_UMI_ASSIGN_LAUNCHES + (deep ? SLR_UMI_ASSIGN_DEEP_LAUNCHES : 0);
    return SLR_OK;
}

int slr_umi_cluster_dev(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets, const int64_t *d_out_offsets, int64_t n_jobs,
                        int64_t n_reads, int ed, const uint8_t *d_member, const int32_t *d_rank, int32_t *d_counts,
                        slr_umi_cluster_rec *d_rec, void *stream)
{
    if (!ctx) return fail(SLR_E_INVALID, "slr_umi_cluster_dev: ctx is NULL");
    if (ed < 0 || ed > 5) return fail(SLR_E_INVALID, "slr_umi_cluster_dev: ed %d outside 0..5", ed);
    if (n_jobs < 0 || n_reads < 0) return fail(SLR_E_INVALID, "slr_umi_cluster_dev: negative size");
    if (n_jobs == 0 || n_reads == 0) return SLR_OK;
    if (!d_matrices || !d_job_offsets || !d_out_offsets || !d_counts || !d_rec) return fail(SLR_E_INVALID, "slr_umi_cluster_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY_WORK(ctx, (cudaStream_t)stream,
                  slr_launch_umi_cluster(d_matrices, (const long long *)d_job_offsets, (const long long *)d_out_offsets, n_jobs, n_reads, ed,
                                         d_member, d_rank, nullptr, d_counts, d_rec, work, (cudaStream_t)stream));
    g_launches += SLR_UMI_CLUSTER_LAUNCHES;
    return SLR_OK;
}

}  // extern "C"

// ---- UMI session: the matrices of one batch stay on the device between the cluster calls ----------------------------------
struct slr_umi_session {
    slr_ctx *ctx = nullptr;
    int64_t n_jobs = 0, n_reads = 0, cells = 0;
    std::vector<int64_t> hoff;                 // host copy of the job offsets (sizes the deep arena of slr_umi_session_assign)
    DevBuf umis, jraw, jtmp, joff, ooff, mat, scr, counts, rec, rank, member, arec, aqv, ascr;
    void release()
    {
        umis.release(); jraw.release(); jtmp.release(); joff.release(); ooff.release(); mat.release(); scr.release(); counts.release();
        rec.release(); rank.release(); member.release(); arec.release(); aqv.release(); ascr.release();
    }
    ~slr_umi_session() { release(); }
};
constexpr size_t SESSION_POOL = 2;             // sessions whose buffers a context keeps for reuse

static void free_session_pool(slr_ctx *c)
{
    std::lock_guard<std::mutex> lk(c->pool_mtx);
    for (slr_umi_session *p : c->session_pool) delete p;
    c->session_pool.clear();
}

extern "C" {

int slr_umi_session_create(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                           slr_umi_session **out)
{
    if (!out) return fail(SLR_E_INVALID, "slr_umi_session_create: out is NULL");
    *out = nullptr;
    int rc = check_umi_args(ctx, stride, umi_len, n_jobs);
    if (rc) return rc;
    if (n_jobs > 0 && (!umis || !job_offsets)) return fail(SLR_E_INVALID, "slr_umi_session_create: NULL buffer");
    int64_t bad = 0;
    for (int64_t k = 0; k < n_jobs; k++) bad |= job_offsets[k + 1] - job_offsets[k];       // the sign bit survives the OR
    if (bad < 0) return fail(SLR_E_INVALID, "job_offsets not monotone");
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::unique_ptr<slr_umi_session> S;
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mtx);
        if (!ctx->session_pool.empty()) { S.reset(ctx->session_pool.back()); ctx->session_pool.pop_back(); }
    }
    if (!S) S.reset(new slr_umi_session());
    S->ctx = ctx; S->n_jobs = n_jobs; S->cells = 0;
    S->hoff.assign(job_offsets, job_offsets + (n_jobs > 0 ? n_jobs + 1 : 0));
    const int64_t r0 = n_jobs > 0 ? job_offsets[0] : 0;
    S->n_reads = n_jobs > 0 ? job_offsets[n_jobs] - r0 : 0;
    if (S->n_reads > 0) {
        Slot *sl = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
        std::lock_guard<std::mutex> lock(sl->mtx);
        cudaStream_t st = sl->stream[0];
        const size_t off_bytes = (size_t)(n_jobs + 1) * 8, m = (size_t)S->n_reads;
        if ((rc = S->umis.reserve(m * stride)) || (rc = S->jraw.reserve(off_bytes)) || (rc = S->jtmp.reserve(slr_umi_rebase_tmp_bytes(n_jobs))) ||
            (rc = S->joff.reserve(off_bytes)) || (rc = S->ooff.reserve(off_bytes)) || (rc = S->scr.reserve(slr_umi_scratch_bytes(S->n_reads))) ||
            (rc = S->counts.reserve(m * 4)) || (rc = S->rec.reserve(m * 16)) || (rc = S->rank.reserve(m * 4)) || (rc = S->member.reserve(m)))
            return rc;
        CUDA_TRY(cudaMemcpyAsync(S->umis.p, umis + r0 * stride, m * stride, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(S->jraw.p, job_offsets, off_bytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(slr_launch_umi_rebase((const long long *)S->jraw.p, n_jobs, r0, (long long *)S->joff.p, (long long *)S->ooff.p, S->jtmp.p, st));
        g_launches += SLR_UMI_REBASE_LAUNCHES;
        long long cells = 0;
        CUDA_TRY(cudaMemcpyAsync(&cells, (const long long *)S->ooff.p + n_jobs, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        S->cells = cells;
        if ((rc = S->mat.reserve((size_t)cells * 4 + 4))) return rc;
        CUDA_TRY(slr_launch_umi_dist((const uint8_t *)S->umis.p, stride, umi_len, (const long long *)S->joff.p, n_jobs, S->n_reads,
                                     (int32_t *)S->mat.p, (const long long *)S->ooff.p, S->scr.p, st));
        g_launches += SLR_UMI_LAUNCHES;
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    *out = S.release();
    return SLR_OK;
}

int slr_umi_session_cluster(slr_umi_session *S, int ed, const uint8_t *member, const int32_t *rank, slr_umi_cluster_rec *rec)
{
    if (!S) return fail(SLR_E_INVALID, "slr_umi_session_cluster: session is NULL");
    if (ed < 0 || ed > 5) return fail(SLR_E_INVALID, "slr_umi_session_cluster: ed %d outside 0..5", ed);
    if (S->n_reads == 0) return SLR_OK;
    if (!rec) return fail(SLR_E_INVALID, "slr_umi_session_cluster: rec is NULL");
    slr_ctx *ctx = S->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *sl = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(sl->mtx);
    cudaStream_t st = sl->stream[0];
    const size_t m = (size_t)S->n_reads;
    if (rank) CUDA_TRY(cudaMemcpyAsync(S->rank.p, rank, m * 4, cudaMemcpyHostToDevice, st));
    if (member) CUDA_TRY(cudaMemcpyAsync(S->member.p, member, m, cudaMemcpyHostToDevice, st));
    CUDA_TRY_WORK(ctx, st, slr_launch_umi_cluster((const int32_t *)S->mat.p, (const long long *)S->joff.p, (const long long *)S->ooff.p, S->n_jobs,
                                                  S->n_reads, ed, member ? (const uint8_t *)S->member.p : nullptr,
                                                  rank ? (const int32_t *)S->rank.p : nullptr, slr_umi_scratch_rowjob(S->scr.p, S->n_reads),
                                                  (int32_t *)S->counts.p, (slr_umi_cluster_rec *)S->rec.p, work, st));
    g_launches += SLR_UMI_CLUSTER_LAUNCHES;
    CUDA_TRY(cudaMemcpyAsync(rec, S->rec.p, m * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SLR_OK;
}

int slr_umi_session_assign(slr_umi_session *S, const slr_umi_assign_params *params, const uint8_t *job_qv01, slr_umi_assign_rec *rec)
{
    if (!S) return fail(SLR_E_INVALID, "slr_umi_session_assign: session is NULL");
    slr_umi_assign_params P;
    int rc = check_assign_params(params, P);
    if (rc) return rc;
    if (S->n_reads == 0) return SLR_OK;
    if (!rec) return fail(SLR_E_INVALID, "slr_umi_session_assign: rec is NULL");
    slr_ctx *ctx = S->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *sl = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(sl->mtx);
    cudaStream_t st = sl->stream[0];
    const size_t m = (size_t)S->n_reads;
    const long long dw = deep_words_of(S->hoff.data(), 0, S->n_jobs, P);
    const size_t scr_bytes = slr_umi_assign_scratch(S->n_jobs, dw);
    if ((rc = S->arec.reserve(m * sizeof(slr_umi_assign_rec))) || (rc = S->aqv.reserve((size_t)S->n_jobs + 1)) || (rc = S->ascr.reserve(scr_bytes)))
        return rc;
    if (job_qv01) CUDA_TRY(cudaMemcpyAsync(S->aqv.p, job_qv01, (size_t)S->n_jobs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(slr_launch_umi_assign((const int32_t *)S->mat.p, (const long long *)S->joff.p, (const long long *)S->ooff.p, S->n_jobs, S->n_reads, P,
                                   job_qv01 ? (const uint8_t *)S->aqv.p : nullptr, slr_umi_scratch_rowjob(S->scr.p, S->n_reads),
                                   (slr_umi_assign_rec *)S->arec.p, S->ascr.p, scr_bytes, st));
    g_launches += SLR_UMI_ASSIGN_LAUNCHES + (dw > 0 ? SLR_UMI_ASSIGN_DEEP_LAUNCHES : 0);
    CUDA_TRY(cudaMemcpyAsync(rec, S->arec.p, m * sizeof(slr_umi_assign_rec), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SLR_OK;
}

int slr_umi_session_matrices(slr_umi_session *S, int32_t *out, int64_t n_cells)
{
    if (!S) return fail(SLR_E_INVALID, "slr_umi_session_matrices: session is NULL");
    if (n_cells != S->cells) return fail(SLR_E_INVALID, "slr_umi_session_matrices: %lld cells asked, the session holds %lld", (long long)n_cells,
                                         (long long)S->cells);
    if (S->cells == 0) return SLR_OK;
    if (!out) return fail(SLR_E_INVALID, "slr_umi_session_matrices: out is NULL");
    CUDA_TRY(cudaSetDevice(S->ctx->device));
    CUDA_TRY(cudaMemcpy(out, S->mat.p, (size_t)S->cells * 4, cudaMemcpyDeviceToHost));
    return SLR_OK;
}

int64_t slr_umi_session_cells(const slr_umi_session *S) { return S ? S->cells : 0; }
int64_t slr_umi_session_reads(const slr_umi_session *S) { return S ? S->n_reads : 0; }
int64_t slr_umi_session_jobs(const slr_umi_session *S) { return S ? S->n_jobs : 0; }

void slr_umi_session_destroy(slr_umi_session *S)
{
    if (!S) return;
    slr_ctx *ctx = S->ctx;
    cudaSetDevice(ctx->device);
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mtx);
        if (ctx->session_pool.size() < SESSION_POOL) { ctx->session_pool.push_back(S); return; }
    }
    delete S;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// S4: Illumina-guided search

struct slr_guided_sets {
    slr_ctx *ctx = nullptr;
    SlrGuidedSetsDev dev;
    void *d_slots = nullptr, *d_groups = nullptr;
    int seq_len = 0;
};

extern "C" {

int slr_guided_sets_create(slr_ctx *ctx, const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups,
                           const uint64_t *all_keys, int64_t n_all, int all_ed, const uint64_t *empty_keys, int64_t n_empty,
                           int empty_ed, int bc_flavour, int seq_len, slr_guided_sets **out)
{
    if (!ctx || !out || n_groups < 0 || (n_groups > 0 && !group_offsets) || n_all < 0 || n_empty < 0)
        return fail(SLR_E_INVALID, "slr_guided_sets_create: bad argument");
    *out = nullptr;
    if (seq_len < 2 || seq_len > 16) return fail(SLR_E_UNSUPPORTED, "sequence length %d not supported by the guided search (2..16)", seq_len);
    if (n_groups > 0x7FFFFFFFLL) return fail(SLR_E_UNSUPPORTED, "more than 2^31 candidate groups");
    for (int64_t g = 0; g < n_groups; g++)
        if (group_offsets[g + 1] < group_offsets[g]) return fail(SLR_E_INVALID, "group_offsets not monotone at %lld", (long long)g);
    if (n_groups > 0 && group_offsets[n_groups] > group_offsets[0] && !group_keys) return fail(SLR_E_INVALID, "group_keys is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    SlrGuidedSetsHost H;
    slr_guided_build(group_keys, group_offsets, n_groups, all_keys, n_all, empty_keys, n_empty, seq_len, H);
    if (H.slots.size() > 0xFFFFFFFFull) return fail(SLR_E_UNSUPPORTED, "candidate tables exceed 2^32 slots");
    slr_guided_sets *s = new slr_guided_sets();
    s->ctx = ctx; s->seq_len = seq_len;
    cudaError_t e = cudaMalloc(&s->d_slots, H.slots.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_groups, (H.groups.size() + 1) * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_slots, H.slots.data(), H.slots.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !H.groups.empty()) e = cudaMemcpy(s->d_groups, H.groups.data(), H.groups.size() * sizeof(uint2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { slr_guided_sets_destroy(s); return fail(SLR_E_CUDA, "slr_guided_sets_create: %s", cudaGetErrorString(e)); }
    s->dev.slots = (const uint32_t *)s->d_slots;
    s->dev.groups = (const uint2 *)s->d_groups;
    s->dev.n_groups = (int)n_groups;
    s->dev.all_set = H.all_set; s->dev.empty_set = H.empty_set;
    s->dev.all_ed = all_ed; s->dev.empty_ed = empty_ed;
    s->dev.bc_flavour = bc_flavour ? 1 : 0;
    *out = s;
    return SLR_OK;
}

void slr_guided_sets_destroy(slr_guided_sets *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_slots); cudaFree(s->d_groups);
    delete s;
}

static int check_guided_args(const slr_ctx *ctx, const slr_guided_sets *s, int plusminus, int post_len, int stride, int slice_len, int64_t n,
                             int raw_cap, int max_ed)
{
    if (!ctx || !s) return fail(SLR_E_INVALID, "slr_guided_match: ctx / sets is NULL");
    if (s->ctx != ctx) return fail(SLR_E_INVALID, "slr_guided_match: the sets belong to another context");
    if (n < 0 || raw_cap < 0) return fail(SLR_E_INVALID, "slr_guided_match: n / raw_cap negative");
    if (plusminus < 0 || plusminus > 4) return fail(SLR_E_UNSUPPORTED, "posplusminus %d not supported (0..4)", plusminus);
    if (slice_len < 1 || slice_len > 32 || stride < slice_len) return fail(SLR_E_UNSUPPORTED, "slice_len %d / stride %d: slices of at most 32 bytes", slice_len, stride);
    if (post_len < 1 || s->seq_len + post_len > 32) return fail(SLR_E_UNSUPPORTED, "post_len %d: need 1 <= post_len and seq_len + post_len <= 32", post_len);
    if (max_ed < 0 || max_ed > SLR_G_MAX_ED) return fail(SLR_E_UNSUPPORTED, "edit distance %d not supported by the guided search (0..%d)", max_ed, SLR_G_MAX_ED);
    if (max_ed + 1 > post_len) return fail(SLR_E_INVALID, "post_len %d shorter than ed + 1 = %d", post_len, max_ed + 1);
    return SLR_OK;
}

int slr_guided_match_dev(slr_ctx *ctx, const slr_guided_sets *s, int plusminus, int post_len, int bailout, const uint8_t *d_slices,
                         int stride, int slice_len, const int32_t *d_anchor, const int32_t *d_group_id, const int32_t *d_ed, int max_ed,
                         int64_t n, slr_guided_result *d_out, slr_guided_hit *d_raw_out, int raw_cap, void *stream)
{
    int rc = check_guided_args(ctx, s, plusminus, post_len, stride, slice_len, n, raw_cap, max_ed);
    if (rc) return rc;
    if (n == 0) return SLR_OK;
    if (!d_slices || !d_anchor || !d_group_id || !d_ed || !d_out) return fail(SLR_E_INVALID, "slr_guided_match_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *sl = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(sl->mtx);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaStreamWaitEvent(st, sl->gvis_free, 0));                  // the slot's visited tables may still serve an earlier launch
    if (slr_guided_vis_bytes(max_ed, nullptr) > sl->gvis.cap) CUDA_TRY(cudaEventSynchronize(sl->gvis_free));   // about to be reallocated
    if ((rc = sl->gvis.reserve(slr_guided_vis_bytes(max_ed, nullptr)))) return rc;
    CUDA_TRY_WORK(ctx, st, slr_launch_guided_match(s->dev, s->seq_len, plusminus, post_len, bailout, d_slices, stride, slice_len, d_anchor,
                                                   d_group_id, d_ed, max_ed, n, d_out, d_raw_out, raw_cap, sl->gvis.p, work, st));
    CUDA_TRY(cudaEventRecord(sl->gvis_free, st));
    g_launches++;
    return SLR_OK;
}

int slr_guided_match(slr_ctx *ctx, const slr_guided_sets *s, int plusminus, int post_len, int bailout, const uint8_t *slices, int stride,
                     int slice_len, const int32_t *anchor, const int32_t *group_id, const int32_t *ed, int64_t n, slr_guided_result *out,
                     slr_guided_hit *raw_out, int raw_cap)
{
    if (n > 0 && (!slices || !anchor || !group_id || !ed || !out)) return fail(SLR_E_INVALID, "slr_guided_match: NULL buffer");
    int max_ed = 0;
    for (int64_t i = 0; i < n; i++) {
        if (ed[i] < 0) return fail(SLR_E_INVALID, "slr_guided_match: ed[%lld] = %d", (long long)i, ed[i]);
        if (ed[i] > max_ed) max_ed = ed[i];
    }
    int rc = check_guided_args(ctx, s, plusminus, post_len, stride, slice_len, n, raw_cap, max_ed);
    if (rc) return rc;
    if (n == 0) return SLR_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Slot *sl = ctx->slots[ctx->next.fetch_add(1) % (unsigned)ctx->n_slots];
    std::lock_guard<std::mutex> lock(sl->mtx);
    SlotDrain drain(sl);
    cudaStream_t st = sl->stream[0];
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventSynchronize(sl->gvis_free));
    const int64_t CHUNK = 1 << 20;
    const size_t m0 = (size_t)(n < CHUNK ? n : CHUNK);
    if ((rc = sl->gsl.reserve(m0 * stride))) return rc;
    if ((rc = sl->ganc.reserve(m0 * 4))) return rc;
    if ((rc = sl->ggid.reserve(m0 * 4))) return rc;
    if ((rc = sl->ged.reserve(m0 * 4))) return rc;
    if ((rc = sl->gout.reserve(m0 * sizeof(slr_guided_result)))) return rc;
    if (raw_out && raw_cap > 0 && (rc = sl->graw.reserve(m0 * raw_cap * sizeof(slr_guided_hit)))) return rc;
    if ((rc = sl->gvis.reserve(slr_guided_vis_bytes(max_ed, nullptr)))) return rc;
    for (int64_t off = 0; off < n; off += CHUNK) {
        const size_t m = (size_t)(n - off < CHUNK ? n - off : CHUNK);
        CUDA_TRY(cudaMemcpyAsync(sl->gsl.p, slices + off * stride, m * stride, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(sl->ganc.p, anchor + off, m * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(sl->ggid.p, group_id + off, m * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(sl->ged.p, ed + off, m * 4, cudaMemcpyHostToDevice, st));
        slr_guided_hit *d_raw = (raw_out && raw_cap > 0) ? (slr_guided_hit *)sl->graw.p : nullptr;
        if (d_raw) CUDA_TRY(cudaMemsetAsync(d_raw, 0, m * raw_cap * sizeof(slr_guided_hit), st));
        CUDA_TRY_WORK(ctx, st, slr_launch_guided_match(s->dev, s->seq_len, plusminus, post_len, bailout, (const uint8_t *)sl->gsl.p, stride,
                                                       slice_len, (const int32_t *)sl->ganc.p, (const int32_t *)sl->ggid.p,
                                                       (const int32_t *)sl->ged.p, max_ed, (long long)m, (slr_guided_result *)sl->gout.p, d_raw,
                                                       raw_cap, sl->gvis.p, work, st));
        g_launches++;
        CUDA_TRY(cudaMemcpyAsync(out + off, sl->gout.p, m * sizeof(slr_guided_result), cudaMemcpyDeviceToHost, st));
        if (d_raw) CUDA_TRY(cudaMemcpyAsync(raw_out + off * raw_cap, d_raw, m * raw_cap * sizeof(slr_guided_hit), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return SLR_OK;
}

// DynamicEditDistances.getmaxED (DynamicEditDistances.java:L93-L98, lambda L94)
int slr_dyn_max_ed(const int64_t *max_candidates, int n_ed, int count, int plusminus, int cap)
{
    if (!max_candidates || n_ed <= 0) return -1;
    const int64_t need = (int64_t)(int32_t)((uint32_t)count * (uint32_t)(2 * plusminus + 1));     // int arithmetic, then i2l
    int best = -1;
    for (int e = 0; e < n_ed; e++)
        if (max_candidates[e] >= need) best = e;                                                   // filter + max by key
    if (best < 0) return -1;
    if (cap >= 0 && best > cap) best = cap;
    return best;
}

}  // extern "C"
