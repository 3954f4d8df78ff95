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
// Kernel 1: |N(a)| per matrix row (lane per row for small jobs, warp per row otherwise).  Kernel 2: one thread per read c, rows l
// walked in batches of 8 loads, column reads coalesced across the lanes of a warp, rows of deep jobs split over 8 warps.  A tie for the maximum is broken by `rank` (the caller's
// iteration rank of key l) or by ascending index when no rank is given; the number of tied entries is reported so
// that the caller can re-evaluate exactly those reads with its own map.
#include "slr_kernels.h"

namespace {

__device__ __forceinline__ long long uc_job_of(const int32_t *__restrict__ rowjob, const long long *__restrict__ joff, long long n_jobs,
                                               long long r)
{
    if (rowjob) return rowjob[r];                     // left behind by the distance kernels
    long long lo = 0, hi = n_jobs;                    // last j with joff[j] <= r
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (joff[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int UC_BATCH = 8;          // matrix cells in flight per thread in the column walk

__device__ __forceinline__ int uc_ed(int32_t packed) { return (int)(int8_t)(packed & 0xFF); }

constexpr int UC_SLICES = 8;         // warps per CTA: row slices per column tile in kernel 2, rows in flight in kernel 1
constexpr int UC_DEEP = 256;         // jobs with at least this many reads have their rows split over the slices in kernel 2
constexpr int UC_LANE_ROW = 8;       // rows of at most this many cells (one 32-byte sector) are summed by a single lane in kernel 1

// Kernel 1.  CTA = 32 consecutive reads; warp 0 looks up the job of every read (one binary search per lane) and sums the one-sector
// rows itself — consecutive rows of a job are contiguous, so the warp still reads one contiguous stretch.  Longer rows are dealt
// round-robin to the UC_SLICES warps, each walked by a whole warp.
__global__ void __launch_bounds__(32 * UC_SLICES) umi_neigh_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                    const long long *__restrict__ ooff, long long n_jobs, long long n_reads,
                                                                    int ed, const uint8_t *__restrict__ member, const int32_t *__restrict__ rowjob,
                                                                    int32_t *__restrict__ counts)
{
    __shared__ long long s_r0[32], s_n[32];
    __shared__ const int32_t *s_row[32];
    __shared__ unsigned s_todo;
    const int x = threadIdx.x, y = threadIdx.y;
    for (long long base = (long long)blockIdx.x * 32; base < n_reads; base += (long long)gridDim.x * 32) {
        if (y == 0) {
            const long long r = base + x;
            const bool live = r < n_reads;
            long long r0 = 0, n = 0;
            const int32_t *row = mat;
            bool in = false;
            if (live) {
                const long long j = uc_job_of(rowjob, joff, n_jobs, r);
                if (j >= 0) {                                // (a read outside every job has no neighbours)
                    r0 = joff[j]; n = joff[j + 1] - r0;
                    row = mat + ooff[j] + (r - r0) * n;
                    in = !member || member[r];
                }
            }
            const bool wide = in && n > UC_LANE_ROW;
            int cnt = 0;
            if (in && !wide) {
                int32_t cell[UC_LANE_ROW];
#pragma unroll
                for (int v = 0; v < UC_LANE_ROW; v++) cell[v] = v < n ? row[v] : 0x7f;       // all loads of the row in flight together
#pragma unroll
                for (int v = 0; v < UC_LANE_ROW; v++) cnt += v < n && (!member || member[r0 + v]) && uc_ed(cell[v]) <= ed;
            }
            if (live && !wide) counts[r] = cnt;
            s_r0[x] = r0; s_n[x] = n; s_row[x] = row;
            const unsigned todo = __ballot_sync(0xffffffffu, wide);
            if (x == 0) s_todo = todo;
        }
        __syncthreads();
        unsigned todo = s_todo;
        for (int k = 0; todo; k++) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            if (k % UC_SLICES != y) continue;
            const long long w_r0 = s_r0[src], w_n = s_n[src];
            const int32_t *w_row = s_row[src];
            int part = 0;
#pragma unroll 8
            for (long long v = x; v < w_n; v += 32) part += (!member || member[w_r0 + v]) && uc_ed(w_row[v]) <= ed;
            part = __reduce_add_sync(0xffffffffu, part);
            if (x == 0) counts[base + src] = part;
        }
        __syncthreads();                                     // the shared job data is rewritten by the next tile
    }
}

struct UcBest { int count, rank, key, ties; };

// running "first maximum": larger |N(l)| wins, equal |N(l)| counts as a tie and the smaller iteration rank stays
__device__ __forceinline__ void uc_merge(UcBest &b, int cl, int rl, int l, int ties)
{
    if (cl > b.count) { b.count = cl; b.key = l; b.ties = ties; b.rank = rl; }
    else if (cl == b.count) {
        b.ties += ties;
        if (rl < b.rank) { b.key = l; b.rank = rl; }
    }
}

// Kernel 2.  CTA = 32 consecutive reads (x) times UC_SLICES warps (y).  Warp 0 resolves the job of every read; for a deep job the
// rows are split over the warps and the partial choices merged through shared memory, otherwise warp 0 walks all rows.
__global__ void __launch_bounds__(32 * UC_SLICES) umi_assign_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                     const long long *__restrict__ ooff, long long n_jobs, long long n_reads,
                                                                     int ed, const int32_t *__restrict__ rank, const int32_t *__restrict__ rowjob,
                                                                     const int32_t *__restrict__ counts, slr_umi_cluster_rec *__restrict__ out)
{
    __shared__ long long s_r0[32], s_n[32], s_oo[32];
    __shared__ UcBest s_part[UC_SLICES][32];
    const int x = threadIdx.x, y = threadIdx.y;
    for (long long base = (long long)blockIdx.x * 32; base < n_reads; base += (long long)gridDim.x * 32) {
        const long long c = base + x;
        if (y == 0) {
            long long r0 = 0, n = 0, oo = 0;
            if (c < n_reads && counts[c] > 1) {              // reads that are no key need no job
                const long long j = uc_job_of(rowjob, joff, n_jobs, c);
                if (j >= 0) { r0 = joff[j]; n = joff[j + 1] - r0; oo = ooff[j]; }
            }
            s_r0[x] = r0; s_n[x] = n; s_oo[x] = oo;
        }
        __syncthreads();
        const long long r0 = s_r0[x], n = s_n[x];
        const bool deep = n >= UC_DEEP;
        UcBest b = {0, 0, -1, 0};
        if (n > 0 && (deep || y == 0)) {
            const int ni = (int)n;                                                  // a job's read count fits an int (its matrix must fit the GPU)
            const int per = deep ? (ni + UC_SLICES - 1) / UC_SLICES : ni;
            const int l_begin = deep ? y * per : 0, l_end = l_begin + per < ni ? l_begin + per : ni;
            const int32_t *p = mat + s_oo[x] + (c - r0) + (long long)l_begin * n;    // matrix[l][c], one row further per step
            const int32_t *pc = counts + r0 + l_begin;
            const int32_t *pr = rank ? rank + r0 : nullptr;
            int l = l_begin;
            for (; l + UC_BATCH <= l_end; l += UC_BATCH, pc += UC_BATCH) {
                // the choice is a running maximum, so the loads of a batch are issued together before any of them is looked at
                int cnt[UC_BATCH];
                int32_t cell[UC_BATCH];
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++) { cnt[k] = pc[k]; cell[k] = *p; p += n; }
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++)
                    if (cnt[k] > 1 && cnt[k] >= b.count && uc_ed(cell[k]) <= ed) uc_merge(b, cnt[k], pr ? pr[l + k] : l + k, l + k, 1);
            }
            if (l < l_end) {                                                        // the last (for small jobs: the only) batch, predicated
                const int rem = l_end - l;
                int cnt[UC_BATCH];
                int32_t cell[UC_BATCH];
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++) {
                    cnt[k] = k < rem ? pc[k] : 0;
                    cell[k] = k < rem ? *p : 0;
                    if (k < rem) p += n;
                }
#pragma unroll
                for (int k = 0; k < UC_BATCH; k++)
                    if (cnt[k] > 1 && cnt[k] >= b.count && uc_ed(cell[k]) <= ed) uc_merge(b, cnt[k], pr ? pr[l + k] : l + k, l + k, 1);
            }
        }
        s_part[y][x] = b;
        __syncthreads();
        if (y == 0 && c < n_reads) {
            if (deep)
                for (int k = 1; k < UC_SLICES; k++) {
                    const UcBest p = s_part[k][x];
                    if (p.key >= 0) uc_merge(b, p.count, p.rank, p.key, p.ties);
                }
            slr_umi_cluster_rec rec;
            rec.n_neighbours = counts[c];
            rec.best_key = b.key; rec.best_count = b.count; rec.n_ties = b.ties;
            out[c] = rec;
        }
        __syncthreads();                                     // s_r0 / s_part are rewritten by the next tile
    }
}

}  // namespace

cudaError_t slr_launch_umi_cluster(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                   long long n_reads, int ed, const uint8_t *d_member, const int32_t *d_rank, const int32_t *d_rowjob,
                                   int32_t *d_counts, slr_umi_cluster_rec *d_out, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long wave = (long long)sms * 8;                         // 8 CTAs of 256 threads per SM
    long long g1 = (n_reads + 31) / 32, g2 = g1;                          // both kernels: a CTA of 8 warps per 32 consecutive reads
    if (g1 > wave) g1 = wave;
    if (g2 > wave) g2 = wave;
    umi_neigh_kernel<<<(unsigned)g1, dim3(32, UC_SLICES), 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_member,
                                                                       d_rowjob, d_counts);
    umi_assign_kernel<<<(unsigned)g2, dim3(32, UC_SLICES), 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_rank, d_rowjob,
                                                                       d_counts, d_out);
    return cudaGetLastError();
}
