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
// Pass 1: |N(a)| per matrix row.  Pass 2: per read c the rows l walked in batches of 8 loads, column reads coalesced across the
// lanes of a warp.  Each pass is a "flat" kernel (every warp owns 32 consecutive reads; jobs below UC_DEEP reads) followed by a
// "deep" kernel (CTA = 32 reads x 8 warps; the rows of a deep job are spread over the warps) that only visits the index range in
// which the flat kernel met deep jobs.  A tie for the maximum is broken by `rank` (the caller's
// iteration rank of key l) or by ascending index when no rank is given; the number of tied entries is reported so
// that the caller can re-evaluate exactly those reads with its own map.
#include "slr_kernels.h"

namespace {

__device__ __forceinline__ long long uc_job_of(const int32_t *__restrict__ rowjob, const long long *__restrict__ joff, long long n_jobs,
                                               long long r)
{
    if (rowjob) return rowjob[r];                     // left behind by the distance kernels (-1: the read is in no job)
    if (r < joff[0] || r >= joff[n_jobs]) return -1;  // padding rows before the first / behind the last job (as umi_rows_kernel marks them)
    long long lo = 0, hi = n_jobs;                    // last j with joff[j] <= r
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (joff[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int UC_BATCH = 8;          // matrix cells in flight per thread in the column walk
constexpr int UC_SLICES = 8;         // warps per CTA of the deep kernels: rows in flight in pass 1, row slices per column tile in pass 2
constexpr int UC_DEEP = 1024;        // jobs with at least this many reads go to the deep kernels
constexpr int UC_LANE_ROW = 8;       // rows of at most this many cells (one 32-byte sector) are summed by a single lane in pass 1
constexpr int UC_DEEP_BIT = 1 << 30; // counts[r]: the read belongs to a deep job (set by the flat kernel of pass 1)
constexpr int UC_CNT_MASK = UC_DEEP_BIT - 1;

__device__ __forceinline__ int uc_ed(int32_t packed) { return (int)(int8_t)(packed & 0xFF); }

struct UcJob { long long r0, n, oo; };

__device__ __forceinline__ UcJob uc_lookup(const int32_t *__restrict__ rowjob, const long long *__restrict__ joff,
                                            const long long *__restrict__ ooff, long long n_jobs, long long r)
{
    UcJob J = {0, 0, 0};
    const long long j = uc_job_of(rowjob, joff, n_jobs, r);
    if (j >= 0) { J.r0 = joff[j]; J.n = joff[j + 1] - J.r0; J.oo = ooff[j]; }     // (a read outside every job has no neighbours)
    return J;
}

// cells [b, n) of a row, at most 32 * NB of them: one predicated batch, all loads issued before the first use
template <int NB>
__device__ __forceinline__ int uc_row_tail(const int32_t *__restrict__ row, long long r0, long long b, long long n, int ed,
                                           const uint8_t *__restrict__ member, int lane)
{
    int32_t cell[NB];
    bool mv[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const long long v = b + 32 * k + lane;
        const bool ok = v < n;
        cell[k] = ok ? row[v] : 0x7f;
        mv[k] = ok && (!member || member[r0 + v]);
    }
    int part = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) part += mv[k] && uc_ed(cell[k]) <= ed;
    return part;
}

// one matrix row walked by a whole warp: full batches of 32 * UC_BATCH cells, then one predicated batch for the rest
__device__ __forceinline__ int uc_row_count(const int32_t *__restrict__ row, long long r0, long long n, int ed,
                                            const uint8_t *__restrict__ member, int lane)
{
    int part = 0;
    long long b = 0;
    for (; b + 32 * UC_BATCH <= n; b += 32 * UC_BATCH) {
        int32_t cell[UC_BATCH];
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++) cell[k] = row[b + 32 * k + lane];
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++) part += (!member || member[r0 + b + 32 * k + lane]) && uc_ed(cell[k]) <= ed;
    }
    if (b < n) part += n - b <= 64 ? uc_row_tail<2>(row, r0, b, n, ed, member, lane) : uc_row_tail<UC_BATCH>(row, r0, b, n, ed, member, lane);
    return __reduce_add_sync(0xffffffffu, part);
}

// Pass 1, flat.  A warp takes 32 consecutive reads, every lane looks up the job of its own read.  One-sector rows are summed by
// their lane (consecutive rows of a job are contiguous, so the warp still reads one contiguous stretch), longer rows by the whole
// warp one after the other; reads of deep jobs are only marked, and the range they span is recorded for the deep kernels.
__global__ void __launch_bounds__(256) umi_neigh_flat(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                       const long long *__restrict__ ooff, long long n_jobs, long long n_reads, int ed,
                                                       const uint8_t *__restrict__ member, const int32_t *__restrict__ rowjob,
                                                       int32_t *__restrict__ counts, unsigned long long *__restrict__ range)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp0 * 32; base < n_reads; base += n_warps * 32) {
        const long long r = base + lane;
        const bool live = r < n_reads;
        UcJob J = {0, 0, 0};
        if (live) J = uc_lookup(rowjob, joff, ooff, n_jobs, r);
        const int32_t *row = mat + J.oo + (r - J.r0) * J.n;
        const bool in = live && J.n > 0 && (!member || member[r]);
        const bool deep = live && J.n >= UC_DEEP;
        const bool wide = in && !deep && J.n > UC_LANE_ROW;
        int cnt = 0;
        if (in && !deep && !wide) {
            int32_t cell[UC_LANE_ROW];
#pragma unroll
            for (int v = 0; v < UC_LANE_ROW; v++) cell[v] = v < J.n ? row[v] : 0x7f;          // all loads of the row in flight together
#pragma unroll
            for (int v = 0; v < UC_LANE_ROW; v++) cnt += v < J.n && (!member || member[J.r0 + v]) && uc_ed(cell[v]) <= ed;
        }
        unsigned todo = __ballot_sync(0xffffffffu, wide);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long w_r0 = __shfl_sync(0xffffffffu, J.r0, src), w_n = __shfl_sync(0xffffffffu, J.n, src);
            const int32_t *w_row = (const int32_t *)__shfl_sync(0xffffffffu, (unsigned long long)row, src);
            const int part = uc_row_count(w_row, w_r0, w_n, ed, member, lane);
            if (lane == src) cnt = part;
        }
        if (live) counts[r] = deep ? UC_DEEP_BIT : cnt;
        const unsigned dm = __ballot_sync(0xffffffffu, deep);
        if (dm && lane == 0) {
            atomicMin(range, (unsigned long long)(base + __ffs(dm) - 1));
            atomicMax(range + 1, (unsigned long long)(base + 32 - __clz(dm)));
        }
    }
}

// Pass 1, deep.  CTA = 32 consecutive reads of the recorded range; warp 0 resolves the jobs of the marked reads, their rows are dealt
// round-robin to the UC_SLICES warps.
__global__ void __launch_bounds__(32 * UC_SLICES) umi_neigh_deep(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                  const long long *__restrict__ ooff, long long n_jobs, long long n_reads,
                                                                  int ed, const uint8_t *__restrict__ member, const int32_t *__restrict__ rowjob,
                                                                  int32_t *__restrict__ counts, const unsigned long long *__restrict__ range)
{
    __shared__ long long s_r0[32], s_n[32];
    __shared__ const int32_t *s_row[32];
    __shared__ unsigned s_todo;
    const long long hi = (long long)range[1];
    if (hi == 0) return;                                     // no deep job in this launch
    const int x = threadIdx.x, y = threadIdx.y;
    for (long long base = ((long long)range[0] & ~31LL) + (long long)blockIdx.x * 32; base < hi; base += (long long)gridDim.x * 32) {
        if (y == 0) {
            const long long r = base + x;
            bool todo = false;
            if (r < n_reads && (counts[r] & UC_DEEP_BIT)) {
                const UcJob J = uc_lookup(rowjob, joff, ooff, n_jobs, r);
                s_r0[x] = J.r0; s_n[x] = J.n; s_row[x] = mat + J.oo + (r - J.r0) * J.n;
                todo = !member || member[r];                 // a marked read outside `indices` keeps |N| = 0
            }
            const unsigned m = __ballot_sync(0xffffffffu, todo);
            if (x == 0) s_todo = m;
        }
        __syncthreads();
        unsigned todo = s_todo;
        for (int k = 0; todo; k++) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            if (k % UC_SLICES != y) continue;
            const int part = uc_row_count(s_row[src], s_r0[src], s_n[src], ed, member, x);
            if (x == 0) counts[base + src] = part | UC_DEEP_BIT;
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

// rows [l_begin, l_end) of column c: p = &matrix[l_begin][c], pc = &counts[r0 + l_begin], pr = rank of the job's keys or NULL
__device__ __forceinline__ void uc_walk(UcBest &b, const int32_t *__restrict__ p, long long n, const int32_t *__restrict__ pc,
                                        const int32_t *__restrict__ pr, int l_begin, int l_end, int ed)
{
    int l = l_begin;
    for (; l + UC_BATCH <= l_end; l += UC_BATCH, pc += UC_BATCH) {
        // the choice is a running maximum, so the loads of a batch are issued together before any of them is looked at
        int cnt[UC_BATCH];
        int32_t cell[UC_BATCH];
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++) { cnt[k] = pc[k] & UC_CNT_MASK; cell[k] = *p; p += n; }
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++)
            if (cnt[k] > 1 && cnt[k] >= b.count && uc_ed(cell[k]) <= ed) uc_merge(b, cnt[k], pr ? pr[l + k] : l + k, l + k, 1);
    }
    if (l < l_end) {                                         // the last (for small jobs: the only) batch, predicated
        const int rem = l_end - l;
        int cnt[UC_BATCH];
        int32_t cell[UC_BATCH];
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++) {
            cnt[k] = k < rem ? pc[k] & UC_CNT_MASK : 0;
            cell[k] = k < rem ? *p : 0;
            if (k < rem) p += n;
        }
#pragma unroll
        for (int k = 0; k < UC_BATCH; k++)
            if (cnt[k] > 1 && cnt[k] >= b.count && uc_ed(cell[k]) <= ed) uc_merge(b, cnt[k], pr ? pr[l + k] : l + k, l + k, 1);
    }
}

// Pass 2, flat.  One thread per read of a job below UC_DEEP reads: all rows of its column.
__global__ void __launch_bounds__(256) umi_assign_flat(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                        const long long *__restrict__ ooff, long long n_jobs, long long n_reads, int ed,
                                                        const int32_t *__restrict__ rank, const int32_t *__restrict__ rowjob,
                                                        const int32_t *__restrict__ counts, slr_umi_cluster_rec *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_reads; c += stride) {
        const int cv = counts[c];
        if (cv & UC_DEEP_BIT) continue;                      // left to umi_assign_deep
        UcBest b = {0, 0, -1, 0};
        if (cv > 1) {                                        // reads that are no key need no job
            const UcJob J = uc_lookup(rowjob, joff, ooff, n_jobs, c);
            uc_walk(b, mat + J.oo + (c - J.r0), J.n, counts + J.r0, rank ? rank + J.r0 : nullptr, 0, (int)J.n, ed);
        }
        slr_umi_cluster_rec rec;
        rec.n_neighbours = cv; rec.best_key = b.key; rec.best_count = b.count; rec.n_ties = b.ties;
        out[c] = rec;
    }
}

// Pass 2, deep.  CTA = 32 consecutive reads (x) times UC_SLICES warps (y) over the recorded range: the rows of the job are split
// over the warps and the partial choices merged through shared memory (the choice is an associative maximum).
__global__ void __launch_bounds__(32 * UC_SLICES) umi_assign_deep(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                   const long long *__restrict__ ooff, long long n_jobs, long long n_reads,
                                                                   int ed, const int32_t *__restrict__ rank, const int32_t *__restrict__ rowjob,
                                                                   const int32_t *__restrict__ counts, slr_umi_cluster_rec *__restrict__ out,
                                                                   const unsigned long long *__restrict__ range)
{
    __shared__ long long s_r0[32], s_n[32], s_oo[32];
    __shared__ UcBest s_part[UC_SLICES][32];
    const long long hi = (long long)range[1];
    if (hi == 0) return;
    const int x = threadIdx.x, y = threadIdx.y;
    for (long long base = ((long long)range[0] & ~31LL) + (long long)blockIdx.x * 32; base < hi; base += (long long)gridDim.x * 32) {
        const long long c = base + x;
        const int cv = c < n_reads ? counts[c] : 0;
        const bool mine = (cv & UC_DEEP_BIT) != 0;
        if (y == 0) {
            UcJob J = {0, 0, 0};
            if (mine && (cv & UC_CNT_MASK) > 1) J = uc_lookup(rowjob, joff, ooff, n_jobs, c);
            s_r0[x] = J.r0; s_n[x] = J.n; s_oo[x] = J.oo;
        }
        __syncthreads();
        const long long r0 = s_r0[x], n = s_n[x];
        UcBest b = {0, 0, -1, 0};
        if (n > 0) {
            const int ni = (int)n;                           // a job's read count fits an int (its matrix must fit the GPU)
            const int per = (ni + UC_SLICES - 1) / UC_SLICES;
            const int l_begin = y * per < ni ? y * per : ni, l_end = l_begin + per < ni ? l_begin + per : ni;
            uc_walk(b, mat + s_oo[x] + (c - r0) + (long long)l_begin * n, n, counts + r0 + l_begin, rank ? rank + r0 : nullptr, l_begin,
                    l_end, ed);
        }
        s_part[y][x] = b;
        __syncthreads();
        if (y == 0 && mine) {
            for (int k = 1; k < UC_SLICES; k++) {
                const UcBest p = s_part[k][x];
                if (p.key >= 0) uc_merge(b, p.count, p.rank, p.key, p.ties);
            }
            slr_umi_cluster_rec rec;
            rec.n_neighbours = cv & UC_CNT_MASK; rec.best_key = b.key; rec.best_count = b.count; rec.n_ties = b.ties;
            out[c] = rec;
        }
        __syncthreads();                                     // s_r0 / s_part are rewritten by the next tile
    }
}

}  // namespace

cudaError_t slr_launch_umi_cluster(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                   long long n_reads, int ed, const uint8_t *d_member, const int32_t *d_rank, const int32_t *d_rowjob,
                                   int32_t *d_counts, slr_umi_cluster_rec *d_out, unsigned long long *d_range, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long wave = (long long)sms * 8;                         // 8 CTAs of 256 threads per SM
    long long g1 = (n_reads + 255) / 256, g2 = g1;                     // flat kernels: pass 1 a warp per 32 reads, pass 2 a thread per read
    if (g1 > wave) g1 = wave;
    if (g2 > wave) g2 = wave;
    cudaMemsetAsync(d_range, 0xff, 8, stream);                         // [first, last + 1) read of a deep job
    cudaMemsetAsync(d_range + 1, 0, 8, stream);
    umi_neigh_flat<<<(unsigned)g1, 256, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_member, d_rowjob, d_counts,
                                                    d_range);
    umi_neigh_deep<<<(unsigned)wave, dim3(32, UC_SLICES), 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_member,
                                                                      d_rowjob, d_counts, d_range);
    umi_assign_flat<<<(unsigned)g2, 256, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_rank, d_rowjob, d_counts,
                                                     d_out);
    umi_assign_deep<<<(unsigned)wave, dim3(32, UC_SLICES), 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, n_jobs, n_reads, ed, d_rank,
                                                                       d_rowjob, d_counts, d_out, d_range);
    return cudaGetLastError();
}
