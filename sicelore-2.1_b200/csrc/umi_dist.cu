// umi_dist.cu — B200 (sm_100a) kernels for the UMI distance matrices.
//
// Replaces ClusteringEditDistanceBase.generateDistanceMatrix{NonParallel,Paralell}
// (F!com/rw/clustering/ClusteringEditDistanceBase.class, ClusteringEditDistanceBase.java:L168-L259) for all
// (cell, region) jobs of a BAM chunk.  The reference walks one matrix row per task (generateOneRow, L234-L238);
// most jobs hold a handful of reads, so rows are far too short to fill a warp.  Here the unit of work is the READ
// PAIR, flattened over all jobs:
//   1. umi_rows_kernel   one thread per read: its job (binary search in the CSR offsets), the number of pairs it
//                        owns as the row read (n-1-i, at least one item so that the diagonal gets written), the four
//                        bit planes of its codes, and a block-level exclusive scan of the item counts;
//   2. umi_scan_blocks   exclusive scan of the per-block totals (one CTA);
//   3. umi_pairs_kernel  persistent CTAs take chunks of 1024 consecutive items; the row prefix of a chunk is staged in
//                        shared memory, every thread finds its (row, column) by binary search there and runs the nine
//                        bit-parallel Levenshtein comparisons entirely in registers (umi_core.cuh,
//                        slr_umi_best9_planes).  Row cells are written coalesced, the mirrored cell gets the transposed
//                        copy exactly like getTransposedEditDistance (L133).
#include "umi_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int ROWS_PER_BLOCK = 1024;           // umi_rows_kernel: reads per CTA (= threads)
constexpr int PAIR_THREADS = 256;
constexpr int ITEMS_PER_THREAD = 4;
constexpr int CHUNK = PAIR_THREADS * ITEMS_PER_THREAD;
constexpr unsigned FULL = 0xFFFFFFFFu;

// codes of one read as four little-endian words (code i in byte i), any stride / alignment
__device__ __forceinline__ void load_codes(const uint8_t *__restrict__ p, int ncodes, bool vec, uint32_t w[4])
{
    if (vec) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0;
        for (int i = 0; i < ncodes; i++) w[i >> 2] |= (uint32_t)__ldg(p + i) << (8 * (i & 3));
    }
    // bytes past the read's codes never take part (the comparisons look at codes 0 .. umi_len+1 only)
}

__global__ void __launch_bounds__(ROWS_PER_BLOCK)
umi_rows_kernel(const uint8_t *__restrict__ umis, int stride, int ncodes, bool vec, const long long *__restrict__ job_offsets,
                long long n_jobs, long long n_reads, int32_t *__restrict__ rowjob, unsigned long long *__restrict__ planes,
                unsigned long long *__restrict__ ploc, unsigned long long *__restrict__ block_tot)
{
    __shared__ unsigned long long warp_tot[ROWS_PER_BLOCK / 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + threadIdx.x;
    unsigned long long items = 0;
    if (row < n_reads) {
        int32_t job = -1;
        items = 1;                                   // rows outside every job: one item that does nothing
        if (row >= __ldg(job_offsets) && row < __ldg(job_offsets + n_jobs)) {
            long long lo = 0, hi = n_jobs;           // last j with job_offsets[j] <= row
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(job_offsets + mid) <= row) lo = mid; else hi = mid;
            }
            const long long j0 = __ldg(job_offsets + lo), n = __ldg(job_offsets + lo + 1) - j0;
            const long long pairs = n - 1 - (row - j0);
            items = (unsigned long long)(pairs > 0 ? pairs : 1);
            job = (int32_t)lo;
        }
        uint32_t w[4], pl[4];
        load_codes(umis + row * (long long)stride, ncodes, vec, w);
        slr_umi_planes(w, pl);
        rowjob[row] = job;
        planes[row] = slr_umi_planes_pack(pl);
    }
    // exclusive scan of `items` over the CTA
    unsigned long long incl = items;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[wib] = incl;
    __syncthreads();
    if (wib == 0) {
        unsigned long long t = warp_tot[lane];
        unsigned long long ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(FULL, ti, d);
            if (lane >= d) ti += up;
        }
        warp_tot[lane] = ti - t;                     // exclusive over warps
        if (lane == 31) block_tot[blockIdx.x] = ti;
    }
    __syncthreads();
    if (row < n_reads) ploc[row] = warp_tot[wib] + incl - items;
}

// in-place exclusive scan of block_tot[0..nb); block_tot[nb] = total number of items
__global__ void __launch_bounds__(1024) umi_scan_blocks(unsigned long long *__restrict__ block_tot, long long nb)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < nb ? block_tot[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_tot[wib] = incl;
        __syncthreads();
        if (wib == 0) {
            const unsigned long long t = warp_tot[lane];
            unsigned long long ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long up = __shfl_up_sync(FULL, ti, d);
                if (lane >= d) ti += up;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (i < nb) block_tot[i] = carry + warp_tot[wib] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[wib] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[nb] = carry_s;
}

// Rebasing of a range of the caller's CSR job offsets on the device: joff[k] = raw[k] - r0 and ooff = exclusive prefix sum of the
// n_k^2 matrix sizes (block-local scan here, block totals by umi_scan_blocks, then umi_rebase_add).  k runs over 0..n_jobs.
__global__ void __launch_bounds__(1024) umi_rebase_local(const long long *__restrict__ raw, long long n_jobs, long long r0,
                                                          long long *__restrict__ joff, unsigned long long *__restrict__ ooff,
                                                          unsigned long long *__restrict__ block_tot)
{
    __shared__ unsigned long long warp_tot[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long k = (long long)blockIdx.x * 1024 + threadIdx.x;
    unsigned long long sq = 0;
    if (k <= n_jobs) joff[k] = raw[k] - r0;
    if (k < n_jobs) {
        const unsigned long long nj = (unsigned long long)(raw[k + 1] - raw[k]);
        sq = nj * nj;
    }
    unsigned long long incl = sq;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[wib] = incl;
    __syncthreads();
    if (wib == 0) {
        const unsigned long long t = warp_tot[lane];
        unsigned long long ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(FULL, ti, d);
            if (lane >= d) ti += up;
        }
        warp_tot[lane] = ti - t;
    }
    __syncthreads();
    const unsigned long long excl = warp_tot[wib] + incl - sq;
    if (k <= n_jobs) ooff[k] = excl;
    if (threadIdx.x == 1023) block_tot[blockIdx.x] = excl + sq;
}

__global__ void __launch_bounds__(1024) umi_rebase_add(unsigned long long *__restrict__ ooff, long long n_jobs,
                                                        const unsigned long long *__restrict__ block_tot)
{
    const long long k = (long long)blockIdx.x * 1024 + threadIdx.x;
    if (k <= n_jobs) ooff[k] += block_tot[blockIdx.x];
}

template <int L>
__global__ void __launch_bounds__(PAIR_THREADS)
umi_pairs_kernel(const uint8_t *__restrict__ umis, int stride, bool vec, const long long *__restrict__ job_offsets,
                 long long n_reads, int32_t *__restrict__ out, const long long *__restrict__ out_offsets,
                 const int32_t *__restrict__ rowjob, const unsigned long long *__restrict__ planes,
                 const unsigned long long *__restrict__ ploc, const unsigned long long *__restrict__ block_off, long long nb)
{
    __shared__ int rel[CHUNK + 1];                   // rel[t] = first item of row r0 + t, relative to the chunk start
    __shared__ long long r0_s;
    __shared__ unsigned long long p0_s;
    const unsigned long long total = block_off[nb];
    const int tid = threadIdx.x;
    for (unsigned long long chunk0 = (unsigned long long)blockIdx.x * CHUNK; chunk0 < total; chunk0 += (unsigned long long)gridDim.x * CHUNK) {
        if (tid == 0) {
            // row that owns item chunk0: last block b with block_off[b] <= chunk0, then last row in it with prefix <= chunk0
            long long lo = 0, hi = nb;
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (block_off[mid] <= chunk0) lo = mid; else hi = mid;
            }
            const unsigned long long boff = block_off[lo];
            long long rl = lo * ROWS_PER_BLOCK, rh = rl + ROWS_PER_BLOCK;
            if (rh > n_reads) rh = n_reads;
            while (rh - rl > 1) {
                const long long mid = (rl + rh) >> 1;
                if (boff + ploc[mid] <= chunk0) rl = mid; else rh = mid;
            }
            r0_s = rl;
            p0_s = boff + ploc[rl];
        }
        __syncthreads();
        const long long r0 = r0_s;
        for (int t = tid + 1; t <= CHUNK; t += PAIR_THREADS) {
            const long long r = r0 + t;
            int v = CHUNK + 1;                       // beyond the chunk
            if (r < n_reads) {
                const unsigned long long d = block_off[r / ROWS_PER_BLOCK] + ploc[r] - chunk0;   // >= 0: rows after r0 start inside or after the chunk
                if (d < (unsigned long long)(CHUNK + 1)) v = (int)d;
            }
            rel[t] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int it = 0; it < ITEMS_PER_THREAD; it++) {
            const int li = it * PAIR_THREADS + tid;
            if (chunk0 + li >= total) break;
            int lo = 0, hi = CHUNK + 1;              // last t with rel[t] <= li (rel[0] = "-inf")
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (rel[mid] <= li) lo = mid; else hi = mid;
            }
            const long long row = r0 + lo;
            const unsigned long long k = lo == 0 ? chunk0 + li - p0_s : (unsigned long long)(li - rel[lo]);
            const int32_t j = rowjob[row];
            if (j < 0) continue;
            const long long j0 = __ldg(job_offsets + j), n = __ldg(job_offsets + j + 1) - j0, i = row - j0;
            int32_t *mat = out + __ldg(out_offsets + j);
            if (k == 0) mat[i * n + i] = slr_umi_equality();
            if (n - 1 - i < 1) continue;
            const long long col = i + 1 + (long long)k;
            uint32_t cw[4];
            load_codes(umis + (j0 + col) * (long long)stride, L + 2, vec, cw);
            const int32_t e = slr_umi_best9_planes<L>(planes[row], cw);
            mat[i * n + col] = e;
            mat[col * n + i] = slr_umi_transpose(e);
        }
        __syncthreads();
    }
}

template <int L>
cudaError_t launch_pairs(const uint8_t *d_umis, int stride, bool vec, const long long *d_job_offsets, long long n_reads, int32_t *d_out,
                         const long long *d_out_offsets, const int32_t *rowjob, const unsigned long long *planes,
                         const unsigned long long *ploc, const unsigned long long *block_off, long long nb, cudaStream_t stream)
{
    static int resident_ctas[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (resident_ctas[dev] == 0) {
        int bps = 0, sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, umi_pairs_kernel<L>, PAIR_THREADS, 0);
        if (e != cudaSuccess) return e;
        resident_ctas[dev] = sms * (bps > 0 ? bps : 1);
    }
    // persistent CTAs, one wave (a multiple of the SM count): the number of items is only known on the device, CTAs
    // without a chunk return at once
    const long long want = resident_ctas[dev];
    umi_pairs_kernel<L><<<(unsigned)want, PAIR_THREADS, 0, stream>>>(d_umis, stride, vec, d_job_offsets, n_reads, d_out, d_out_offsets,
                                                                    rowjob, planes, ploc, block_off, nb);
    return cudaGetLastError();
}

}  // namespace

size_t slr_umi_rebase_tmp_bytes(long long n_jobs) { return (size_t)((n_jobs + 1 + 1023) / 1024 + 1) * 8; }

cudaError_t slr_launch_umi_rebase(const long long *d_raw, long long n_jobs, long long r0, long long *d_joff, long long *d_ooff, void *d_tmp,
                                  cudaStream_t stream)
{
    const long long nb = (n_jobs + 1 + 1023) / 1024;
    unsigned long long *block_tot = reinterpret_cast<unsigned long long *>(d_tmp);
    umi_rebase_local<<<(unsigned)nb, 1024, 0, stream>>>(d_raw, n_jobs, r0, d_joff, reinterpret_cast<unsigned long long *>(d_ooff), block_tot);
    umi_scan_blocks<<<1, 1024, 0, stream>>>(block_tot, nb);
    umi_rebase_add<<<(unsigned)nb, 1024, 0, stream>>>(reinterpret_cast<unsigned long long *>(d_ooff), n_jobs, block_tot);
    return cudaGetLastError();
}

size_t slr_umi_scratch_bytes(long long n_reads)
{
    const size_t nb = (size_t)((n_reads + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
    return (size_t)n_reads * 20 + (nb + 1) * 8 + 64;
}

const int32_t *slr_umi_scratch_rowjob(const void *d_scratch, long long n_reads)
{
    const long long nb = (n_reads + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;     // same layout as in slr_launch_umi_dist
    return reinterpret_cast<const int32_t *>(reinterpret_cast<const unsigned long long *>(d_scratch) + 2 * n_reads + nb + 1);
}

cudaError_t slr_launch_umi_dist(const uint8_t *d_umis, int stride, int umi_len, const long long *d_job_offsets, long long n_jobs,
                                long long n_reads, int32_t *d_out, const long long *d_out_offsets, void *d_scratch, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    const long long nb = (n_reads + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    // scratch layout: planes u64[m] | ploc u64[m] | block_off u64[nb + 1] | rowjob i32[m]
    unsigned long long *planes = reinterpret_cast<unsigned long long *>(d_scratch);
    unsigned long long *ploc = planes + n_reads;
    unsigned long long *block_off = ploc + n_reads;
    int32_t *rowjob = reinterpret_cast<int32_t *>(block_off + nb + 1);
    const bool vec = stride % 16 == 0 && (reinterpret_cast<uintptr_t>(d_umis) & 15u) == 0;
    umi_rows_kernel<<<(unsigned)nb, ROWS_PER_BLOCK, 0, stream>>>(d_umis, stride, umi_len + 2, vec, d_job_offsets, n_jobs, n_reads, rowjob,
                                                                planes, ploc, block_off);
    umi_scan_blocks<<<1, 1024, 0, stream>>>(block_off, nb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
#define SLR_UMI_CASE(LL)                                                                                                    \
    case LL:                                                                                                                \
        return launch_pairs<LL>(d_umis, stride, vec, d_job_offsets, n_reads, d_out, d_out_offsets, rowjob, planes, ploc, block_off, nb, stream);
    switch (umi_len) {
        SLR_UMI_CASE(1) SLR_UMI_CASE(2) SLR_UMI_CASE(3) SLR_UMI_CASE(4) SLR_UMI_CASE(5) SLR_UMI_CASE(6) SLR_UMI_CASE(7)
        SLR_UMI_CASE(8) SLR_UMI_CASE(9) SLR_UMI_CASE(10) SLR_UMI_CASE(11) SLR_UMI_CASE(12) SLR_UMI_CASE(13) SLR_UMI_CASE(14)
    default: return cudaErrorInvalidValue;
    }
#undef SLR_UMI_CASE
}
