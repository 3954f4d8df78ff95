// umi_dist.cu — B200 (sm_100a) kernel for the UMI distance matrices.
//
// Replaces ClusteringEditDistanceBase.generateDistanceMatrix{NonParallel,Paralell}
// (F!com/rw/clustering/ClusteringEditDistanceBase.class, ClusteringEditDistanceBase.java:L168-L259) for all
// (cell, region) jobs of a BAM chunk: one warp per matrix row (the reference's generateOneRow, L234-L238,
// one CompletableFuture per row).  The row read's three shifted UMI windows become Peq tables in shared
// memory once per warp; every lane then takes one column read as the text and runs the nine bit-parallel
// Levenshtein comparisons (umi_core.cuh).  Row cells are written coalesced, the mirrored column cells get the
// transposed copy exactly like getTransposedEditDistance (L133).
#include "umi_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int UMI_WARPS_PER_BLOCK = 8;

__global__ void __launch_bounds__(UMI_WARPS_PER_BLOCK * 32)
umi_dist_kernel(const uint8_t *__restrict__ umis, int stride, int umi_len, const long long *__restrict__ job_offsets,
                long long n_jobs, long long n_reads, int32_t *__restrict__ out, const long long *__restrict__ out_offsets)
{
    __shared__ uint32_t peq_s[UMI_WARPS_PER_BLOCK][48];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * UMI_WARPS_PER_BLOCK + wib;
    if (row >= n_reads) return;
    // job of this row: last j with job_offsets[j] <= row (uniform binary search)
    long long lo = 0, hi = n_jobs;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(job_offsets + mid) <= row) lo = mid; else hi = mid;
    }
    const long long j0 = __ldg(job_offsets + lo), n = __ldg(job_offsets + lo + 1) - j0;
    const long long i = row - j0;
    int32_t *mat = out + __ldg(out_offsets + lo);

    const int ncodes = umi_len + 2;
    const unsigned long long rowp = slr_umi_pack(umis + row * (long long)stride, ncodes);
    uint32_t *peq = peq_s[wib];
    for (int e = lane; e < 48; e += 32) peq[e] = slr_umi_peq_entry(rowp, umi_len, e >> 4, (uint32_t)(e & 15));
    __syncwarp();

    for (long long v = i + lane; v < n; v += 32) {
        if (v == i) { mat[i * n + i] = slr_umi_equality(); continue; }
        const unsigned long long colp = slr_umi_pack(umis + (j0 + v) * (long long)stride, ncodes);
        const int32_t e = slr_umi_best9(peq, umi_len, colp);
        mat[i * n + v] = e;
        mat[v * n + i] = slr_umi_transpose(e);
    }
}

}  // namespace

cudaError_t slr_launch_umi_dist(const uint8_t *d_umis, int stride, int umi_len, const long long *d_job_offsets, long long n_jobs,
                                long long n_reads, int32_t *d_out, const long long *d_out_offsets, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    const long long blocks = (n_reads + UMI_WARPS_PER_BLOCK - 1) / UMI_WARPS_PER_BLOCK;
    umi_dist_kernel<<<(unsigned)blocks, UMI_WARPS_PER_BLOCK * 32, 0, stream>>>(d_umis, stride, umi_len, d_job_offsets, n_jobs, n_reads,
                                                                             d_out, d_out_offsets);
    return cudaGetLastError();
}
