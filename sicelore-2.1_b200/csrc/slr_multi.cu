// slr_multi.cu — several B200s behind ONE caller (the reference is a single JVM: WorkerReadscanner.java:L186-L204 runs every worker in
// one process, so the multi-GPU path must be reachable through the C ABI without a launcher).
//
//   slr_multi_create            one slr_ctx per device, peer access enabled between all pairs (NVLink / NVSwitch)
//   slr_multi_bc_table_create   the search set replicated on every device
//   slr_multi_bc_assign / _exact   contiguous read shards, one host thread per device driving the single-device pipeline, results positional
//   slr_multi_bc_counts_read    the per-barcode x ED counters of all replicas summed ON THE DEVICE: device 0 reads its peers' counter arrays
//                               over NVLink (P2P loads) in one kernel, then one copy to the host — the BarcodesAssigned.tsv merge
//   slr_multi_umi_*             whole (cell, region) jobs dealt to the devices in contiguous runs balanced by their n^2 cost: a job is never
//                               cut, so no cross-device merge exists on this path (a launcher that shards the read STREAM by index has one:
//                               see UmiShardMerger in the Python mirror / bench.py)
#include <cstdio>
#include <string>
#include <thread>
#include <vector>
#include "slr_kernels.h"

extern "C" int slr_multi_fail(int code, const char *msg);      // slr_api.cu: sets the thread-local error message

struct slr_multi {
    std::vector<slr_ctx *> ctx;
    std::vector<int> dev;
    bool peer = false;                                         // every device can read every other device's memory
};
struct slr_multi_table {
    slr_multi *m = nullptr;
    std::vector<slr_bc_table *> t;
    int64_t n = 0;
    unsigned long long **d_ptrs = nullptr;                     // on device 0: the counter arrays of all replicas
    unsigned long long *d_sum = nullptr;                       // on device 0: 3 n sums
};

namespace {

__global__ void __launch_bounds__(256) counts_reduce_kernel(unsigned long long *const *__restrict__ ptrs, int n_dev, long long n,
                                                            unsigned long long *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long s = 0;
        for (int d = 0; d < n_dev; d++) s += ptrs[d][i];       // d > 0: a load over NVLink from the peer's HBM
        out[i] = s;
    }
}

// contiguous shares of n units over k workers
inline void split_even(int64_t n, int k, std::vector<int64_t> &cut)
{
    cut.assign((size_t)k + 1, 0);
    for (int i = 0; i <= k; i++) cut[(size_t)i] = n * i / k;
}

// run f(i) for every device on its own host thread; first non-zero return code wins (its message is re-raised on the caller's thread)
template <class F>
int for_each_device(slr_multi *m, F f)
{
    const int k = (int)m->ctx.size();
    std::vector<int> rc((size_t)k, 0);
    std::vector<std::string> msg((size_t)k);
    if (k == 1) {
        rc[0] = f(0);
        return rc[0];
    }
    std::vector<std::thread> th;
    for (int i = 0; i < k; i++)
        th.emplace_back([&, i] {
            rc[(size_t)i] = f(i);
            if (rc[(size_t)i]) msg[(size_t)i] = slr_last_error();
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < k; i++)
        if (rc[(size_t)i]) return slr_multi_fail(rc[(size_t)i], msg[(size_t)i].c_str());
    return SLR_OK;
}

}  // namespace

extern "C" {

int slr_multi_create(int n_devices, const int *device_ids, int n_streams, slr_multi **out)
{
    if (!out) return slr_multi_fail(SLR_E_INVALID, "slr_multi_create: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return slr_multi_fail(SLR_E_NODEVICE, "no CUDA device available; libsicelore_gpu has no CPU fallback");
    if (n_devices <= 0) n_devices = count;                     // 0 = every visible device
    if (n_devices > count && !device_ids) return slr_multi_fail(SLR_E_INVALID, "slr_multi_create: more devices asked than visible");
    slr_multi *m = new slr_multi();
    for (int i = 0; i < n_devices; i++) {
        const int d = device_ids ? device_ids[i] : i;
        slr_ctx *c = nullptr;
        const int rc = slr_ctx_create(d, n_streams, &c);
        if (rc) { slr_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
        m->dev.push_back(d);
    }
    m->peer = true;
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < n_devices; j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[(size_t)i], m->dev[(size_t)j]);
            if (!can) { m->peer = false; continue; }
            cudaSetDevice(m->dev[(size_t)i]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[(size_t)j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer = false;
            cudaGetLastError();
        }
    *out = m;
    return SLR_OK;
}

void slr_multi_destroy(slr_multi *m)
{
    if (!m) return;
    for (slr_ctx *c : m->ctx) slr_ctx_destroy(c);
    delete m;
}

int slr_multi_n_devices(const slr_multi *m) { return m ? (int)m->ctx.size() : 0; }
slr_ctx *slr_multi_ctx(slr_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
int slr_multi_peer_access(const slr_multi *m) { return m && m->peer ? 1 : 0; }

int slr_multi_bc_table_create(slr_multi *m, const uint64_t *barcodes2bit, const int32_t *rank, int64_t n, int bc_len, slr_multi_table **out)
{
    if (!m || !out) return slr_multi_fail(SLR_E_INVALID, "slr_multi_bc_table_create: NULL argument");
    *out = nullptr;
    slr_multi_table *T = new slr_multi_table();
    T->m = m; T->n = n;
    T->t.assign(m->ctx.size(), nullptr);
    const int rc = for_each_device(m, [&](int i) { return slr_bc_table_create(m->ctx[(size_t)i], barcodes2bit, rank, n, bc_len, &T->t[(size_t)i]); });
    if (rc) { slr_multi_bc_table_destroy(T); return rc; }
    *out = T;
    return SLR_OK;
}

void slr_multi_bc_table_destroy(slr_multi_table *T)
{
    if (!T) return;
    for (slr_bc_table *t : T->t) slr_bc_table_destroy(t);
    if (T->m && !T->m->dev.empty()) {
        cudaSetDevice(T->m->dev[0]);
        cudaFree(T->d_ptrs);
        cudaFree(T->d_sum);
    }
    delete T;
}

slr_bc_table *slr_multi_bc_table_replica(slr_multi_table *T, int i) { return (T && i >= 0 && i < (int)T->t.size()) ? T->t[(size_t)i] : nullptr; }

static int multi_bc(slr_multi *m, const slr_multi_table *T, int exact, int ed_max, int plusminus, int three_prime, const uint8_t *slices, int stride,
                    int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out)
{
    if (!m || !T || T->m != m) return slr_multi_fail(SLR_E_INVALID, "slr_multi_bc_assign: NULL argument / the table belongs to another device set");
    if (n < 0) return slr_multi_fail(SLR_E_INVALID, "n < 0");
    if (n == 0) return SLR_OK;
    if (!slices || !anchor || !out) return slr_multi_fail(SLR_E_INVALID, "slr_multi_bc_assign: NULL buffer");
    std::vector<int64_t> cut;
    split_even(n, (int)m->ctx.size(), cut);
    return for_each_device(m, [&](int i) {
        const int64_t a = cut[(size_t)i], k = cut[(size_t)i + 1] - a;
        if (k == 0) return (int)SLR_OK;
        const uint8_t *sl = slices + a * stride;
        const int32_t *ln = lens ? lens + a : nullptr;
        return exact ? slr_bc_exact(m->ctx[(size_t)i], T->t[(size_t)i], three_prime, sl, stride, slice_len, ln, anchor + a, k, out + a)
                     : slr_bc_assign(m->ctx[(size_t)i], T->t[(size_t)i], ed_max, plusminus, three_prime, sl, stride, slice_len, ln, anchor + a, k,
                                     out + a);
    });
}

int slr_multi_bc_assign(slr_multi *m, const slr_multi_table *T, int ed_max, int plusminus, int three_prime, const uint8_t *slices, int stride,
                        int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out)
{
    return multi_bc(m, T, 0, ed_max, plusminus, three_prime, slices, stride, slice_len, lens, anchor, n, out);
}

int slr_multi_bc_exact(slr_multi *m, const slr_multi_table *T, int three_prime, const uint8_t *slices, int stride, int slice_len,
                       const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out)
{
    return multi_bc(m, T, 1, 0, 0, three_prime, slices, stride, slice_len, lens, anchor, n, out);
}

int slr_multi_bc_counts_read(slr_multi *m, slr_multi_table *T, int64_t *counts_out)
{
    if (!m || !T || T->m != m || !counts_out) return slr_multi_fail(SLR_E_INVALID, "slr_multi_bc_counts_read: NULL argument");
    const int k = (int)m->ctx.size();
    const long long n3 = (long long)T->n * 3;
    if (n3 == 0) return SLR_OK;
    if (k == 1) return slr_bc_counts_read(m->ctx[0], T->t[0], counts_out);
    // every replica's launches have to be complete: the host-pointer calls return synchronised; _dev callers synchronise their streams
    std::vector<int64_t *> ptrs((size_t)k, nullptr);
    for (int i = 0; i < k; i++) {
        int64_t ne = 0;
        const int rc = slr_bc_counts_device(T->t[(size_t)i], &ptrs[(size_t)i], &ne);
        if (rc) return rc;
    }
    cudaError_t e = cudaSetDevice(m->dev[0]);
    if (e == cudaSuccess && !T->d_sum) e = cudaMalloc((void **)&T->d_sum, (size_t)n3 * 8);
    if (e == cudaSuccess && !T->d_ptrs) e = cudaMalloc((void **)&T->d_ptrs, (size_t)k * sizeof(void *));
    std::vector<unsigned long long *> staged;                  // without peer access: copies of the peers' arrays on device 0
    if (e == cudaSuccess && !m->peer) {
        for (int i = 1; i < k && e == cudaSuccess; i++) {
            unsigned long long *p = nullptr;
            e = cudaMalloc((void **)&p, (size_t)n3 * 8);
            if (e == cudaSuccess) { staged.push_back(p); e = cudaMemcpyPeer(p, m->dev[0], ptrs[(size_t)i], m->dev[(size_t)i], (size_t)n3 * 8); }
            ptrs[(size_t)i] = (int64_t *)p;
        }
    }
    if (e == cudaSuccess) e = cudaMemcpy(T->d_ptrs, ptrs.data(), (size_t)k * sizeof(void *), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        for (int i = 0; i < k; i++) { cudaSetDevice(m->dev[(size_t)i]); cudaDeviceSynchronize(); }
        cudaSetDevice(m->dev[0]);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->dev[0]);
        long long g = (n3 + 255) / 256;
        if (g > (long long)sms * 8) g = (long long)sms * 8;
        counts_reduce_kernel<<<(unsigned)g, 256>>>(T->d_ptrs, k, n3, T->d_sum);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(counts_out, T->d_sum, (size_t)n3 * 8, cudaMemcpyDeviceToHost);
    for (unsigned long long *p : staged) cudaFree(p);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "slr_multi_bc_counts_read: %s", cudaGetErrorString(e));
        return slr_multi_fail(SLR_E_CUDA, buf);
    }
    return SLR_OK;
}

int slr_multi_bc_counts_reset(slr_multi *m, slr_multi_table *T)
{
    if (!m || !T || T->m != m) return slr_multi_fail(SLR_E_INVALID, "slr_multi_bc_counts_reset: NULL argument");
    return for_each_device(m, [&](int i) { return slr_bc_counts_reset(m->ctx[(size_t)i], T->t[(size_t)i]); });
}

// ---- UMI jobs: contiguous runs of whole jobs per device, balanced by the jobs' n^2 (matrix cells = work) ------------------------------
static void split_jobs(const int64_t *job_offsets, int64_t n_jobs, int k, std::vector<int64_t> &cut)
{
    cut.assign((size_t)k + 1, n_jobs);
    cut[0] = 0;
    double total = 0;
    for (int64_t j = 0; j < n_jobs; j++) { const double nj = (double)(job_offsets[j + 1] - job_offsets[j]); total += nj * nj + 8 * nj; }
    double acc = 0;
    int w = 1;
    for (int64_t j = 0; j < n_jobs && w < k; j++) {
        const double nj = (double)(job_offsets[j + 1] - job_offsets[j]);
        acc += nj * nj + 8 * nj;
        while (w < k && acc >= total * w / k) cut[(size_t)w++] = j + 1;
    }
}

int slr_multi_umi_assign(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                         const slr_umi_assign_params *params, const uint8_t *job_qv01, slr_umi_assign_rec *rec)
{
    if (!m) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_assign: NULL argument");
    if (n_jobs < 0) return slr_multi_fail(SLR_E_INVALID, "n_jobs < 0");
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !rec) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_assign: NULL buffer");
    std::vector<int64_t> cut;
    split_jobs(job_offsets, n_jobs, (int)m->ctx.size(), cut);
    // the single-device call addresses reads through the caller's own job offsets, so every device gets the same base pointers
    return for_each_device(m, [&](int i) {
        const int64_t a = cut[(size_t)i], k = cut[(size_t)i + 1] - a;
        if (k <= 0) return (int)SLR_OK;
        return slr_umi_assign(m->ctx[(size_t)i], umis, stride, umi_len, job_offsets + a, k, params, job_qv01 ? job_qv01 + a : nullptr, nullptr, nullptr,
                              rec);
    });
}

int slr_multi_umi_cluster(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int ed,
                          const uint8_t *member, const int32_t *rank, slr_umi_cluster_rec *rec)
{
    if (!m) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_cluster: NULL argument");
    if (n_jobs < 0) return slr_multi_fail(SLR_E_INVALID, "n_jobs < 0");
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !rec) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_cluster: NULL buffer");
    std::vector<int64_t> cut;
    split_jobs(job_offsets, n_jobs, (int)m->ctx.size(), cut);
    return for_each_device(m, [&](int i) {
        const int64_t a = cut[(size_t)i], k = cut[(size_t)i + 1] - a;
        if (k <= 0) return (int)SLR_OK;
        return slr_umi_cluster(m->ctx[(size_t)i], umis, stride, umi_len, job_offsets + a, k, ed, member, rank, nullptr, nullptr, rec);
    });
}

int slr_multi_umi_dist(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int32_t *out,
                       const int64_t *out_offsets)
{
    if (!m) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_dist: NULL argument");
    if (n_jobs < 0) return slr_multi_fail(SLR_E_INVALID, "n_jobs < 0");
    if (n_jobs == 0) return SLR_OK;
    if (!umis || !job_offsets || !out || !out_offsets) return slr_multi_fail(SLR_E_INVALID, "slr_multi_umi_dist: NULL buffer");
    std::vector<int64_t> cut;
    split_jobs(job_offsets, n_jobs, (int)m->ctx.size(), cut);
    return for_each_device(m, [&](int i) {
        const int64_t a = cut[(size_t)i], k = cut[(size_t)i + 1] - a;
        if (k <= 0) return (int)SLR_OK;
        return slr_umi_dist(m->ctx[(size_t)i], umis, stride, umi_len, job_offsets + a, k, out, out_offsets + a);
    });
}

}  // extern "C"
