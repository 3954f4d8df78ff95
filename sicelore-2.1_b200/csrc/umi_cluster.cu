// umi_cluster.cu — neighbour sets and "best cluster key" of ClusterOne_MyClustering.clusterLocal on the packed matrices
//
// Reference (F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class, ClusterOne_MyClustering.java:L175-L219):
//   possibleClusters : a -> N(a) = { v in indices : getED(matrix[a][v]) <= ed }, kept when |N(a)| > 1          (L179-L185)
//   for every key c  : the entry l with c in N(l) and the largest |N(l)| (Stream.max: first maximum in the iteration
//                      order of the fastutil map)                                                               (L190-L196)
//   clusters         : keys grouped by that entry                                                               (L199, L219)
// Both steps are O(n^2) matrix reads per (cell, region) job.  getED = (byte)(packed & 0xFFFFFF)
// (ClusteringEditDistanceBase$BestEditDistance.java:L382).
//
// Kernel 1: one warp per matrix row -> |N(a)|.  Kernel 2: one thread per read c, rows l walked in ascending order in
// batches of 8 loads, column reads coalesced across the lanes of a warp.  A tie for the maximum is broken by `rank` (the caller's
// iteration rank of key l) or by ascending index when no rank is given; the number of tied entries is reported so
// that the caller can re-evaluate exactly those reads with its own map.
#include "slr_kernels.h"

namespace {

__device__ __forceinline__ long long uc_job_of(const long long *__restrict__ joff, long long n_jobs, long long r)
{
    long long lo = 0, hi = n_jobs;                    // last j with joff[j] <= r
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (joff[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int UC_BATCH = 8;          // matrix cells in flight per thread in the column walk

__device__ __forceinline__ int uc_ed(int32_t packed) { return (int)(int8_t)(packed & 0xFF); }

__global__ void __launch_bounds__(256) umi_neigh_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                         const long long *__restrict__ ooff, long long n_jobs, long long n_reads, int ed,
                                                         const uint8_t *__restrict__ member, int32_t *__restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < n_reads; r += n_warps) {
        const long long j = uc_job_of(joff, n_jobs, r);
        const long long r0 = joff[j], n = joff[j + 1] - r0;
        int cnt = 0;
        if (!member || member[r]) {
            const int32_t *row = mat + ooff[j] + (r - r0) * n;
#pragma unroll 8
            for (long long v = lane; v < n; v += 32)
                cnt += (!member || member[r0 + v]) && uc_ed(row[v]) <= ed;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) counts[r] = cnt;
    }
}

__global__ void __launch_bounds__(256) umi_assign_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                          const long long *__restrict__ ooff, long long n_jobs, long long n_reads, int ed,
                                                          const int32_t *__restrict__ rank, const int32_t *__restrict__ counts,
                                                          slr_umi_cluster_rec *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_reads; c += stride) {
        slr_umi_cluster_rec rec;
        rec.n_neighbours = counts[c];
        rec.best_key = -1; rec.best_count = 0; rec.n_ties = 0;
        if (rec.n_neighbours > 1) {
            const long long j = uc_job_of(joff, n_jobs, c);
            const long long r0 = joff[j], n = joff[j + 1] - r0;
            const int32_t *col = mat + ooff[j] + (c - r0);
            int best_rank = 0;
            for (long long l0 = 0; l0 < n; l0 += UC_BATCH) {
                // the choice is a running maximum, so the loads of a batch are issued together before any of them is looked at
                int cnt[UC_BATCH];
                int32_t cell[UC_BATCH];
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++) {
                    const long long l = l0 + k < n ? l0 + k : n - 1;
                    cnt[k] = counts[r0 + l];
                    cell[k] = col[l * n];
                }
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++) {
                    const long long l = l0 + k;
                    const int cl = cnt[k];
                    if (l >= n || cl <= 1 || cl < rec.best_count || uc_ed(cell[k]) > ed) continue;
                    const int rl = rank ? rank[r0 + l] : (int)l;
                    if (cl > rec.best_count) { rec.best_count = cl; rec.best_key = (int32_t)l; rec.n_ties = 1; best_rank = rl; }
                    else {
                        rec.n_ties++;
                        if (rl < best_rank) { rec.best_key = (int32_t)l; best_rank = rl; }
                    }
                }
            }
        }
        out[c] = rec;
    }
}

}  // namespace

cudaError_t slr_launch_umi_cluster(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                   long long n_reads, int ed, const uint8_t *d_member, const int32_t *d_rank, int32_t *d_counts,
                                   slr_umi_cluster_rec *d_out, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long wave = (long long)sms * 8;                         // 8 CTAs of 256 threads per SM
    long long g1 = (n_reads * 32 + 255) / 256, g2 = (n_reads + 255) / 256;
    if (g1 > wave) g1 = wave;
    if (g2 > wave) g2 = wave;
    umi_neigh_kernel<<<(unsigned)g1, 256, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_member, d_counts);
    umi_assign_kernel<<<(unsigned)g2, 256, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_rank, d_counts, d_out);
    return cudaGetLastError();
}
