// umi_assign_deep.cu — clustering and UMI assignment of the LARGE (cell, region) jobs (more than 100 reads) on the packed matrices in HBM.
//
// Replaces ClusterOne_MyClustering.call (F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class, ClusterOne_MyClustering.java:
// L59-L166) — the clusterer UmiClustering$Submitter picks for every job above 100 reads (UmiClustering.java:L240) — as a whole:
//   clusterLocal over all reads            L175-L219: N(a) = reads within umi_completelinkclusteringED of a, kept when |N(a)| > 1; every such read
//                                          joins the entry with the LARGEST neighbour set that contains it (Stream.max keeps the first of equal
//                                          maxima in the iteration order of the fastutil Int2ObjectOpenHashMap); groups -> HashSet<Set<Integer>>
//   depth rule                             L77-L84: size * foldDepthBelowMaxDiscardForClustering > largest cluster, else flagDontUMIassignRecords
//   OneUmiCluster.setClusterCenter         F!com/rw/clustering/OneUmiCluster.class (OneUmiCluster.java:L49-L65)
//   removeOffCenter                        L60-L65, L102: members farther than ED from the centre leave the cluster (OneUmiCluster.removeEntries,
//                                          L114-L119: AbstractCollection.removeAll through the fastutil iterator, then the centre is chosen again)
//   clusterLocal over the unclustered      L104-L112: reads never clustered + the removed ones, clusters of more than one read are appended
//   per read                               L116-L164 -> ClusterOneBase.setSamflagsAndStatsForClustered (ClusterOneBase.java:L118-L168)
// Every stream of that class is parallel above 30 reads; these kernels (like oracle/slr_oracle_assign.c, which they are tested against) have the
// SEQUENTIAL semantics, i.e. the result of a JVM with one worker thread, and reproduce the container orders that reach it: fastutil open
// addressing (insertion, growth, iterator-driven removal), java.util.HashSet bucket order, ConcurrentHashMap bins and their transfer, the
// HashSet of the clusters (hash = sum of the members).  SLR_UA_TIE_UNPIN marks a job in which a read could choose between largest neighbour
// sets that are not the same set, or a bin reached the treeify threshold.
//
// One TEAM per job: one CTA (jobs up to SLR_UA_DEEP_SMALL reads), a thread-block cluster of 8 CTAs (up to SLR_UA_DEEP_MEDIUM), or — one giant job
// after the other — the whole GPU as a cooperative grid.  The O(n^2) passes over the matrix (neighbour
// counts, entry choice, sums of squared distances, U2) are spread over the team's warps with coalesced row reads; the hash-table emulations are
// inherently sequential and run on the team's first thread between team barriers.  All working arrays live in the caller's scratch.
#include <cooperative_groups.h>
#include <mutex>
#include "slr_kernels.h"

namespace cg = cooperative_groups;

namespace {

constexpr unsigned FULLM = 0xFFFFFFFFu;
constexpr int DEEP_THREADS = 512;

__device__ __forceinline__ int dp_ed(int32_t p) { return (int)(int8_t)(p & 0xFF); }
__device__ __forceinline__ int dp_pos1_offset(int32_t p) { return (p & 0x08000000) ? -1 : ((p & 0x10000000) ? 0 : ((p & 0x20000000) ? 1 : 0)); }
__device__ __forceinline__ int dp_pos2_code(int32_t p) { return (p & 0x01000000) ? 0 : ((p & 0x02000000) ? 1 : ((p & 0x04000000) ? 2 : 1)); }
__device__ __forceinline__ uint32_t dp_mix(int k) { const uint32_t h = (uint32_t)k * 0x9E3779B9u; return h ^ (h >> 16); }
__device__ __forceinline__ uint32_t dp_spread(uint32_t h) { return h ^ (h >> 16); }
__device__ __forceinline__ unsigned long long dp_sig(int v)
{
    unsigned long long h = (unsigned long long)(v + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return h;
}
__device__ __forceinline__ int dp_jdk_cap(int size) { int cap = 16; while (size > cap / 4 * 3) cap <<= 1; return cap; }
__device__ __forceinline__ int dp_max_fill(int n) { const int c = n / 4 * 3; return c < n - 1 ? c : n - 1; }     // n is a power of two >= 32
__device__ __forceinline__ int dp_array_size(int expected)
{
    const long long need = ((long long)expected * 4 + 2) / 3;     // ceil(expected / .75)
    long long n = 2;
    while (n < need) n <<= 1;
    return (int)n;
}

// header words of a job's scratch
enum { H_NK = 0, H_NG, H_NCL, H_NU, H_NREM, H_FLAG, H_TOTAL, H_NBIG, H_Q0, H_WORDS = 16 };
constexpr int SMALL_K = 24;                    // clusters up to this size never grow fastutil's initial 32-slot table: one thread, local arrays

struct DeepW {
    int *hdr, *cnt, *chosen, *keys, *tmp, *first, *corder, *perm, *mem, *it, *pos_cl, *sumsq, *gm, *pscr;
    int *cl_beg, *cl_len0, *cl_len, *cl_center, *cl_nvict, *cl_dirty, *cl_offmean, *cl_nf, *cl_entry, *rem_off, *big_list;
    int *clid, *victim, *idx, *inU, *tabA, *tabB, *wrapped, *chmA, *chmB, *chm_nxt;
    int *ck, *firstpos, *gcount, *isfirst, *grp_q;
    unsigned *ht, *chm_h, *gsum, *adj;        // adj: n rows of (n + 31) / 32 words, bit j of row a = ED(matrix[a][j]) <= ed
    unsigned long long *sig;
};
__device__ inline void carve(int *W, int n, DeepW &w)
{
    const int s = (n + 3) & ~1;                                    // even: the 64-bit views (sig, best64 in gm, ksig in pscr) stay 8-byte aligned
    w.sig = reinterpret_cast<unsigned long long *>(W);             // 8-byte aligned: every arena offset is even
    int *p = W + 2 * s;
    w.hdr = p; p += 64;
    auto take = [&](int words) { int *q = p; p += words; return q; };
    w.cnt = take(s); w.chosen = take(s); w.keys = take(s); w.tmp = take(s); w.first = take(s); w.corder = take(s); w.perm = take(s);          // 7
    w.mem = take(2 * s); w.it = take(2 * s); w.pos_cl = take(2 * s); w.sumsq = take(2 * s); w.gm = take(2 * s); w.pscr = take(2 * s);        // 12
    w.cl_beg = take(s); w.cl_len0 = take(s); w.cl_len = take(s); w.cl_center = take(s); w.cl_nvict = take(s); w.cl_dirty = take(s);
    w.cl_offmean = take(s); w.cl_nf = take(s); w.cl_entry = take(s); w.rem_off = take(s); w.big_list = take(s);                                 // 11
    w.clid = take(s); w.victim = take(s); w.idx = take(2 * s); w.inU = take(s); w.wrapped = take(s); w.chm_nxt = take(s);                      // 7
    w.ck = take(s); w.firstpos = take(s); w.gcount = take(s); w.isfirst = take(s); w.grp_q = take(s);                                         // 5
    w.ht = reinterpret_cast<unsigned *>(take(s)); w.chm_h = reinterpret_cast<unsigned *>(take(s)); w.gsum = reinterpret_cast<unsigned *>(take(s));   // 3
    // 2 (sig) + 45 of the 52 (n + 2)-word units slr_umi_assign_deep_words() grants
    w.tabA = take(3 * n + 64); w.tabB = take(3 * n + 64); w.chmA = take(4 * n + 64); w.chmB = take(4 * n + 64);
    w.adj = reinterpret_cast<unsigned *>(p);                       // n * ((n + 31) / 32) words, the last piece of the job's arena
}

// ---- team = 1 CTA or a cluster of CS CTAs ---------------------------------------------------------------------------------------------------
template <int CS>                              // CS = 1: one CTA, CS = 8: a cluster of 8 CTAs, CS = 0: the whole (cooperative) grid
struct Team {
    __device__ static __forceinline__ void sync()
    {
        if (CS == 1) __syncthreads();
        else if (CS == 0) { __threadfence(); cg::this_grid().sync(); }
        else { __threadfence(); cg::this_cluster().sync(); }
    }
    __device__ static __forceinline__ int rank() { return CS == 1 ? 0 : (CS == 0 ? (int)blockIdx.x : (int)cg::this_cluster().block_rank()); }
    __device__ static __forceinline__ int tid() { return rank() * DEEP_THREADS + (int)threadIdx.x; }
    __device__ static __forceinline__ int size() { return (CS == 0 ? (int)gridDim.x : CS) * DEEP_THREADS; }
};

// ---- sequential container emulations (one thread) ---------------------------------------------------------------------------------------------
// fastutil open-addressing table filled in the given order (keys distinct); returns the table size, *tab_out = the table, *has_zero
__device__ int fu_fill(const int *in, int k, int *bufA, int *bufB, int **tab_out, int *has_zero)
{
    int n = 32, size = 0, hz = 0;
    int *tab = bufA, *alt = bufB;
    for (int i = 0; i < n; i++) tab[i] = 0;
    for (int i0 = 0; i0 < k; i0 += 8) {
        int kk[8];                                                      // the input loads of 8 inserts in flight together (one thread: latency is all)
#pragma unroll
        for (int u = 0; u < 8; u++) kk[u] = i0 + u < k ? in[i0 + u] : -1;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int key = kk[u];
            if (i0 + u >= k) break;
            if (key == 0) hz = 1;
            else {
                int pos = (int)(dp_mix(key) & (uint32_t)(n - 1));
                while (tab[pos] != 0) pos = (pos + 1) & (n - 1);
                tab[pos] = key;
            }
            if (size++ >= dp_max_fill(n)) {
                const int nn = dp_array_size(size + 1);
                for (int j = 0; j < nn; j++) alt[j] = 0;
                for (int j = n - 1; j >= 0; j--)
                    if (tab[j] != 0) {
                        int pos = (int)(dp_mix(tab[j]) & (uint32_t)(nn - 1));
                        while (alt[pos] != 0) pos = (pos + 1) & (nn - 1);
                        alt[pos] = tab[j];
                    }
                int *t = tab; tab = alt; alt = t; n = nn;
            }
        }
    }
    *tab_out = tab; *has_zero = hz;
    return n;
}
__device__ int fu_iter(const int *tab, int n, int has_zero, int *out)
{
    int o = 0;
    if (has_zero) out[o++] = 0;
    for (int j = n - 1; j >= 0; j--) if (tab[j] != 0) out[o++] = tab[j];
    return o;
}
__device__ void fu_shift(int *key, int mask, int pos, int *wrapped, int *n_wrapped)
{
    for (;;) {
        const int last = pos;
        int curr;
        pos = (pos + 1) & mask;
        for (;;) {
            if ((curr = key[pos]) == 0) { key[last] = 0; return; }
            const int slot = (int)(dp_mix(curr) & (uint32_t)mask);
            if (last <= pos ? (last >= slot || slot > pos) : (last >= slot && slot > pos)) break;
            pos = (pos + 1) & mask;
        }
        if (wrapped && pos < last) wrapped[(*n_wrapped)++] = curr;
        key[last] = curr;
    }
}
// IntOpenHashSet.remove(int) incl. the shrink rule; the table may move to `alt`
__device__ void fu_remove(int **tabp, int **altp, int *np, int *sizep, int *hzp, int k)
{
    int *tab = *tabp, n = *np;
    if (k == 0) { if (!*hzp) return; *hzp = 0; (*sizep)--; }
    else {
        int pos = (int)(dp_mix(k) & (uint32_t)(n - 1));
        while (tab[pos] != k) { if (tab[pos] == 0) return; pos = (pos + 1) & (n - 1); }
        (*sizep)--;
        fu_shift(tab, n - 1, pos, nullptr, nullptr);
    }
    if (n > 32 && *sizep < dp_max_fill(n) / 4 && n > 16) {       // n > minN (32) && size < maxFill / 4 && n > DEFAULT_INITIAL_SIZE
        const int nn = n / 2;
        int *alt = *altp;
        for (int j = 0; j < nn; j++) alt[j] = 0;
        for (int j = n - 1; j >= 0; j--)
            if (tab[j] != 0) {
                int pos = (int)(dp_mix(tab[j]) & (uint32_t)(nn - 1));
                while (alt[pos] != 0) pos = (pos + 1) & (nn - 1);
                alt[pos] = tab[j];
            }
        *tabp = alt; *altp = tab; *np = nn;
    }
}
// java.util.AbstractCollection.removeAll driven by the set's iterator; victim[x] != 0 marks the elements to drop
__device__ void fu_remove_all(int **tabp, int **altp, int *np, int *sizep, int *hzp, const int *victim, int *wrapped)
{
    int pos = *np, c = *sizep, must_null = *hzp, n_wrapped = 0;
    while (c != 0) {
        c--;
        if (must_null) { must_null = 0; if (victim[0]) { *hzp = 0; (*sizep)--; } continue; }
        for (;;) {
            if (--pos < 0) { const int cur = wrapped[-pos - 1]; if (victim[cur]) fu_remove(tabp, altp, np, sizep, hzp, cur); break; }
            const int cur = (*tabp)[pos];
            if (cur != 0) { if (victim[cur]) { fu_shift(*tabp, *np - 1, pos, wrapped, &n_wrapped); (*sizep)--; } break; }
        }
    }
}
// iteration order of a java.util.HashSet filled in the given order: perm = indices of the elements in iteration order; cnt: cap + 1 words
__device__ void jdk_order(const unsigned *hash, int k, int *perm, int *cnt, int *long_bin)
{
    const int cap = dp_jdk_cap(k);
    for (int b = 0; b <= cap; b++) cnt[b] = 0;
    for (int i = 0; i < k; i++) cnt[(dp_spread(hash[i]) & (uint32_t)(cap - 1)) + 1]++;
    for (int b = 0; b < cap; b++) { if (cnt[b + 1] >= 9) *long_bin = 1; cnt[b + 1] += cnt[b]; }
    for (int i = 0; i < k; i++) perm[cnt[dp_spread(hash[i]) & (uint32_t)(cap - 1)]++] = i;
}
// iteration order of a ConcurrentHashMap<Integer, ?> filled by one thread (computeIfAbsent) in the given order; headA / headB: 4 K + 64 words
__device__ void chm_order(const int *keys, int K, int *out, int *headA, int *headB, int *nxt, unsigned *hsh, int *long_bin)
{
    int cap = 16, sc = 12, count = 0;
    int *head = headA, *alt = headB;
    for (int i = 0; i < cap; i++) head[i] = -1;
    for (int t = 0; t < K; t++) {
        const unsigned h = dp_spread((unsigned)keys[t]) & 0x7FFFFFFFu;
        const int b = (int)(h & (unsigned)(cap - 1));
        hsh[t] = h; nxt[t] = -1;
        int len = 0, last = -1;
        for (int p = head[b]; p >= 0; p = nxt[p]) { last = p; len++; }
        if (len >= 8) *long_bin = 1;
        if (last < 0) head[b] = t; else nxt[last] = t;
        count++;
        while (count >= sc) {                                           // addCount -> transfer: last run kept, the nodes before it prepended
            for (int i = 0; i < 2 * cap; i++) alt[i] = -1;
            for (int i = 0; i < cap; i++) {
                const int f = head[i];
                if (f < 0) continue;
                unsigned run_bit = hsh[f] & (unsigned)cap;
                int last_run = f;
                for (int p = nxt[f]; p >= 0; p = nxt[p]) { const unsigned bb = hsh[p] & (unsigned)cap; if (bb != run_bit) { run_bit = bb; last_run = p; } }
                int ln = -1, hn = -1;
                if (run_bit == 0) ln = last_run; else hn = last_run;
                for (int p = f; p != last_run;) {
                    const int pn = nxt[p];
                    if ((hsh[p] & (unsigned)cap) == 0) { nxt[p] = ln; ln = p; } else { nxt[p] = hn; hn = p; }
                    p = pn;
                }
                alt[i] = ln; alt[i + cap] = hn;
            }
            int *tsw = head; head = alt; alt = tsw;
            cap *= 2; sc = cap - (cap >> 2);
        }
    }
    int o = 0;
    for (int i = 0; i < cap; i++) for (int p = head[i]; p >= 0; p = nxt[p]) out[o++] = keys[p];
}

// ---- parallel passes ---------------------------------------------------------------------------------------------------------------------------
// neighbour counts and set signatures of idx[0 .. L) against the reads marked in inU (NULL = all)
constexpr int DU = 8;                          // independent loads a lane keeps in flight in the row passes
template <int CS>
__device__ void pass_counts(const int32_t *__restrict__ M, int n, int ed, const int *idx, int L, const int *inU, DeepW &w)
{
    const int lane = threadIdx.x & 31, warp = Team<CS>::tid() >> 5, n_warps = Team<CS>::size() >> 5;
    const int W = (n + 31) >> 5;
    for (int i = warp; i < L; i += n_warps) {
        const int a = idx ? idx[i] : i;
        unsigned *arow = w.adj + (size_t)a * W;
        int c = 0;
        unsigned long long s = 0;
        if (!inU) {
            // round 1: one pass over the packed row; the threshold test is kept as one bit per cell — the entry choice (two sweeps over
            // keys x reads) and round 2 then move 1/32 of the bytes, and the bit matrix of a 20 000-read job (50 MB) stays in L2
            const int32_t *row = M + (size_t)a * n;
            for (int j0 = 0; j0 < n; j0 += 32 * DU) {
                int32_t v[DU];
#pragma unroll
                for (int u = 0; u < DU; u++) { const int j = j0 + u * 32 + lane; v[u] = j < n ? __ldg(row + j) : 0x7F; }
#pragma unroll
                for (int u = 0; u < DU; u++) {
                    const bool hit = dp_ed(v[u]) <= ed;
                    const unsigned bits = __ballot_sync(FULLM, hit);
                    if (j0 + u * 32 < n && lane == 0) arow[(j0 >> 5) + u] = bits;
                    if (hit) { c++; s += dp_sig(j0 + u * 32 + lane); }
                }
            }
        } else {
            for (int wd = lane; wd < W; wd += 32) {
                unsigned bits = arow[wd];
                while (bits) {
                    const int j = (wd << 5) + __ffs((int)bits) - 1;
                    bits &= bits - 1;
                    if (inU[j]) { c++; s += dp_sig(j); }
                }
            }
        }
        c = __reduce_add_sync(FULLM, c);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULLM, s, o);
        if (lane == 0) { w.cnt[a] = c; w.sig[a] = s; }
    }
}
// every key c joins the first entry (in key order) with the largest neighbour set that contains c.  The key range is cut into slices so that the
// whole team works on a job of any size: slice maxima meet in a 64-bit atomicMax on (|N(e)|, -position), a second sweep looks for entries tied
// with the winner whose neighbour set is a different one.  kcnt / ksig: |N| and set signature of the keys in key order.
template <int CS>
__device__ void pass_choose_prepare(DeepW &w, unsigned long long *best64)
{
    const int nk = w.hdr[H_NK];
    for (int p = Team<CS>::tid(); p < nk; p += Team<CS>::size()) {
        const int e = w.keys[p];
        w.ck[p] = w.cnt[e];                                             // (ck is rewritten by groups_scan afterwards)
        reinterpret_cast<unsigned long long *>(w.pscr)[p] = w.sig[e];
        best64[e] = 0;
    }
}
template <int CS, bool SECOND>
__device__ void pass_choose_sweep(const int32_t *__restrict__ M, int n, int ed, const int *inU, DeepW &w, unsigned long long *best64)
{
    const int nk = w.hdr[H_NK], T = Team<CS>::size();
    const unsigned long long *ksig = reinterpret_cast<const unsigned long long *>(w.pscr);
    const int n_slices = T / n > 0 ? T / n : 1, per = (nk + n_slices - 1) / n_slices;
    int harmful = 0;
    for (int t = Team<CS>::tid(); t < n * n_slices; t += T) {
        const int c = t % n, sl = t / n;
        if ((inU && !inU[c]) || w.cnt[c] <= 1) continue;
        const int p_lo = sl * per, p_hi = min(nk, p_lo + per);
        int bestcnt = -1, bestp = -1;
        unsigned long long bestsig = 0;
        if (SECOND) {
            const unsigned long long b = best64[c];
            bestcnt = (int)(b >> 32); bestp = (int)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFu));
            bestsig = ksig[bestp];
        }
        const int W = (n + 31) >> 5, cw = c >> 5, cb = c & 31;
        for (int p0 = p_lo; p0 < p_hi; p0 += DU) {
            int e[DU];
            unsigned v[DU];
#pragma unroll
            for (int u = 0; u < DU; u++) {
                e[u] = p0 + u < p_hi ? w.keys[p0 + u] : -1;
                v[u] = e[u] >= 0 ? w.adj[(size_t)e[u] * W + cw] : 0u;     // matrix[e][c] <= ed (L195: entries whose neighbour set contains c)
            }
#pragma unroll
            for (int u = 0; u < DU; u++) {
                if (e[u] < 0 || !((v[u] >> cb) & 1u)) continue;
                const int ce = w.ck[p0 + u];
                if (SECOND) { if (ce == bestcnt && ksig[p0 + u] != bestsig) harmful = 1; }
                else if (ce > bestcnt) { bestcnt = ce; bestp = p0 + u; }
            }
        }
        if (!SECOND && bestp >= 0) atomicMax(&best64[c], ((unsigned long long)(unsigned)bestcnt << 32) | (0xFFFFFFFFu - (unsigned)bestp));
        if (SECOND && sl == 0) w.chosen[c] = w.keys[bestp];
    }
    if (harmful) atomicOr(&w.hdr[H_FLAG], 1);
}
// sum of squared distances of every member of a dirty cluster to the other members
template <int CS>
__device__ void pass_sumsq(const int32_t *__restrict__ M, int n, DeepW &w)
{
    const int lane = threadIdx.x & 31, warp = Team<CS>::tid() >> 5, n_warps = Team<CS>::size() >> 5;
    const int total = w.hdr[H_TOTAL];
    for (int pos = warp; pos < total; pos += n_warps) {
        const int c = w.pos_cl[pos];
        const int m = w.cl_len[c];
        if (!w.cl_dirty[c] || pos - w.cl_beg[c] >= m || m <= 2) continue;
        const int a = w.it[pos];
        const int32_t *row = M + (size_t)a * n;
        int s = 0;
        if ((long long)m * 8 >= n) {
            for (int j0 = 0; j0 < n; j0 += 32 * DU) {
                int32_t v[DU];
                int cl[DU];
#pragma unroll
                for (int u = 0; u < DU; u++) {
                    const int j = j0 + u * 32 + lane;
                    v[u] = j < n ? __ldg(row + j) : 0;
                    cl[u] = j < n ? w.clid[j] : -2;
                }
#pragma unroll
                for (int u = 0; u < DU; u++)
                    if (cl[u] == c && j0 + u * 32 + lane != a) { const int e = dp_ed(v[u]); s += e * e; }
            }
        } else {
            const int *seg = w.it + w.cl_beg[c];
            for (int q0 = 0; q0 < m; q0 += 32 * DU) {
                int x[DU];
                int32_t v[DU];
#pragma unroll
                for (int u = 0; u < DU; u++) {
                    const int q = q0 + u * 32 + lane;
                    x[u] = q < m ? seg[q] : a;
                    v[u] = __ldg(row + x[u]);
                }
#pragma unroll
                for (int u = 0; u < DU; u++)
                    if (x[u] != a) { const int e = dp_ed(v[u]); s += e * e; }
            }
        }
        s = __reduce_add_sync(FULLM, s);
        if (lane == 0) w.sumsq[pos] = s;
    }
}
template <int CS>
__device__ void pass_centers(int qv01, DeepW &w)
{
    const int n_cl = w.hdr[H_NCL];
    for (int c = Team<CS>::tid(); c < n_cl; c += Team<CS>::size()) {
        if (!w.cl_dirty[c]) continue;
        const int m = w.cl_len[c], b = w.cl_beg[c];
        int center = w.it[b];
        if (m == 2) center = qv01 ? w.it[b] : w.it[b + 1];
        else if (m > 2) {
            int best = w.sumsq[b];
            for (int q = 1; q < m; q++) if (w.sumsq[b + q] < best) { best = w.sumsq[b + q]; center = w.it[b + q]; }
        }
        w.cl_center[c] = center;
    }
}
template <int CS>
__device__ void pass_victims(const int32_t *__restrict__ M, int n, int ed, DeepW &w)
{
    const int total = w.hdr[H_TOTAL];
    for (int pos = Team<CS>::tid(); pos < total; pos += Team<CS>::size()) {
        const int c = w.pos_cl[pos];
        if (pos - w.cl_beg[c] >= w.cl_len[c]) continue;
        const int a = w.it[pos];
        if (dp_ed(M[(size_t)a * n + w.cl_center[c]]) > ed) { w.victim[a] = 1; atomicAdd(&w.cl_nvict[c], 1); }
    }
}

// phase time stamps (ns, %globaltimer) in the job's header words 32 ... 63: read by tools/perf_deep.py, free otherwise
#define DEEP_STAMP(w, i) do { if (Team<CS>::tid() == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
                                                       reinterpret_cast<unsigned long long *>((w).hdr + 32)[i] = t_; } } while (0)

// ---- table buffers: the team's shared memory when the table fits, the job's global scratch otherwise ----------------------------------------------
struct TabBufs { int *A, *B, *cnt; };          // A >= final table, B >= half of it, cnt >= JDK capacity + 1 words
template <int SM_INTS>
__device__ __forceinline__ bool fits_smem(int k)
{
    constexpr int A = SM_INTS / 3 * 2;         // 2/3 + 1/3 split: A is a power of two, B = A / 2
    return dp_array_size(k + 1) <= A && dp_jdk_cap(k) + 1 <= SM_INTS;
}
template <int SM_INTS>
__device__ __forceinline__ TabBufs bufs_for(int k, int *smem, DeepW &w)
{
    TabBufs t;
    if (fits_smem<SM_INTS>(k)) { t.A = smem; t.B = smem + SM_INTS / 3 * 2; t.cnt = smem; }
    else { t.A = w.tabA; t.B = w.tabB; t.cnt = w.tabA; }
    return t;
}
// fastutil table of k keys with the FINAL table in bufs.A (the growth steps alternate between the two buffers)
__device__ int fu_fill_into_a(const int *in, int k, const TabBufs &tb, int **tab_out, int *has_zero)
{
    int T = 32, g = 0;
    for (int sz = 1; sz <= k; sz++) if (sz - 1 >= dp_max_fill(T)) { T = dp_array_size(sz + 1); g++; }
    return (g & 1) ? fu_fill(in, k, tb.B, tb.A, tab_out, has_zero) : fu_fill(in, k, tb.A, tb.B, tab_out, has_zero);
}

// ---- the same table, filled by a whole CTA --------------------------------------------------------------------------------------------------
// Linear probing is first-come-first-served: a key ends in the first slot at or after its home that no EARLIER key holds.  That fixed point is
// reached from any schedule in which a slot always goes to the earliest key that asked for it, so all keys probe at once: atomicMin of the
// insertion time per slot, losers step on, until nobody moves (as many rounds as the longest probe run).  Growth is replayed stage by stage: a
// rehash re-inserts the old table from its last slot to its first, i.e. the old keys get the times T_old - 1 - slot, the keys that follow keep
// their insertion order behind them.  own: the table (>= final size words), pos / tim: k words each.  Returns the table size; the table then
// holds the keys (0 = empty).  Collective over the CTA.
__device__ int fu_fill_par(const int *in, int k, int *own, int *pos, int *tim, int *has_zero)
{
    __shared__ int s_moved, s_zero;
    const int tid = threadIdx.x;
    if (tid == 0) s_zero = -1;
    __syncthreads();
    for (int i = tid; i < k; i += DEEP_THREADS) if (in[i] == 0) s_zero = i;
    __syncthreads();
    const int zero_idx = s_zero;
    int T = 32, done = 0, Tprev = 0;
    for (;;) {
        const int mf = dp_max_fill(T), n_end = k < mf + 1 ? k : mf + 1, mask = T - 1;
        for (int i = tid; i < n_end; i += DEEP_THREADS) {
            if (i == zero_idx) continue;
            tim[i] = i < done ? Tprev - 1 - pos[i] : Tprev + i;
            pos[i] = (int)(dp_mix(in[i]) & (uint32_t)mask);
        }
        for (int j = tid; j < T; j += DEEP_THREADS) own[j] = 0x7FFFFFFF;
        __syncthreads();
        for (;;) {
            if (tid == 0) s_moved = 0;
            __syncthreads();
            int moved = 0;
            for (int i = tid; i < n_end; i += DEEP_THREADS) {
                if (i == zero_idx) continue;
                const int t = tim[i];
                int p = pos[i];
                if (own[p] == t) continue;                              // holds its slot (so far)
                for (;;) {                                              // a slot claimed by an earlier key stays with an earlier key: step on at once
                    const int old = atomicMin(&own[p], t);
                    if (old >= t) break;
                    p = (p + 1) & mask;
                }
                pos[i] = p;
                moved = 1;
            }
            if (moved) s_moved = 1;
            __syncthreads();
            const int again = s_moved;
            __syncthreads();
            if (!again) break;
        }
        done = n_end;
        if (n_end == k && k <= mf) break;                               // the last insert did not trigger a rehash
        Tprev = T;
        T = dp_array_size(n_end + 1);                                   // if (size++ >= maxFill) rehash(arraySize(size + 1, f))
    }
    for (int j = tid; j < T; j += DEEP_THREADS) own[j] = 0;
    __syncthreads();
    for (int i = tid; i < k; i += DEEP_THREADS) if (i != zero_idx) own[pos[i]] = in[i];
    __syncthreads();
    *has_zero = zero_idx >= 0;
    return T;
}
// iteration order of a table (key 0 first, then the slots from the last to the first), by a whole CTA; returns the number of keys
__device__ int fu_iter_par(const int *tab, int T, int has_zero, int *out)
{
    __shared__ int s_cnt[DEEP_THREADS + 1];
    const int tid = threadIdx.x, per = (T + DEEP_THREADS - 1) / DEEP_THREADS;
    const int hi = T - 1 - tid * per, lo = hi - per + 1 < 0 ? 0 : hi - per + 1;      // thread 0 owns the last slots
    int c = 0;
    for (int j = hi; j >= lo; j--) c += tab[j] != 0;
    s_cnt[tid] = c;
    __syncthreads();
    if (tid == 0) {
        int acc = has_zero ? 1 : 0;
        if (has_zero) out[0] = 0;
        for (int t = 0; t < DEEP_THREADS; t++) { const int v = s_cnt[t]; s_cnt[t] = acc; acc += v; }
        s_cnt[DEEP_THREADS] = acc;
    }
    __syncthreads();
    int o = s_cnt[tid];
    for (int j = hi; j >= lo; j--) if (tab[j] != 0) out[o++] = tab[j];
    const int total = s_cnt[DEEP_THREADS];
    __syncthreads();
    return total;
}
// ordered compaction by a whole CTA: out = (val(i) for i in 0 .. n - 1 if pred(i)), returns the count; every thread owns a contiguous piece
template <typename P, typename V>
__device__ int compact_par(int n, P pred, V val, int *out)
{
    __shared__ int s_c[DEEP_THREADS + 1];
    const int tid = threadIdx.x, per = (n + DEEP_THREADS - 1) / DEEP_THREADS, lo = tid * per, hi = lo + per < n ? lo + per : n;
    int c = 0;
    for (int i = lo; i < hi; i++) c += pred(i) ? 1 : 0;
    s_c[tid] = c;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int t = 0; t < DEEP_THREADS; t++) { const int v = s_c[t]; s_c[t] = acc; acc += v; }
        s_c[DEEP_THREADS] = acc;
    }
    __syncthreads();
    int o = s_c[tid];
    for (int i = lo; i < hi; i++) if (pred(i)) out[o++] = val(i);
    const int total = s_c[DEEP_THREADS];
    __syncthreads();
    return total;
}
// ConcurrentHashMap order by a CTA: the 16 bins of the initial table never mix (a transfer sends bin i to i and i + capacity), so 16 threads
// replay one sixteenth of the map each, in the keys' global order (the transfers are triggered by the global count).  Tables, links and hashes
// live in shared memory: F + F / 2 + 2 K words (F = final capacity); returns false when that does not fit.  out_idx: K words of scratch.
template <int SM_INTS>
__device__ bool chm_order_par(const int *keys, int K, int *out, int *out_idx, int *smem, int *long_bin)
{
    int F = 16, doublings = 0;
    while (K >= F - (F >> 2)) { F <<= 1; doublings++; }
    if (F + F / 2 + 2 * K > SM_INTS) return false;
    int *A = smem, *B = smem + F, *nxt = smem + F + F / 2;
    unsigned *hsh = reinterpret_cast<unsigned *>(nxt + K);
    __shared__ int s_long;
    const int tid = threadIdx.x;
    for (int t = tid; t < K; t += DEEP_THREADS) { hsh[t] = dp_spread((unsigned)keys[t]) & 0x7FFFFFFFu; nxt[t] = -1; }
    if (tid == 0) s_long = 0;
    __syncthreads();
    if ((tid & 31) == 0 && (tid >> 5) < 16) {                           // 16 workers, one per warp (lanes of one warp would serialise)
        const int r = tid >> 5;
        int cap = 16, sc = 12, lb = 0;
        int *head = (doublings & 1) ? B : A, *alt = (doublings & 1) ? A : B;      // the final table lands in A
        head[r] = -1;
        for (int t = 0; t < K; t++) {
            const unsigned h = hsh[t];
            if ((int)(h & 15u) == r) {
                const int b = (int)(h & (unsigned)(cap - 1));
                int len = 0, last = -1;
                for (int p = head[b]; p >= 0; p = nxt[p]) { last = p; len++; }
                if (len >= 8) lb = 1;
                if (last < 0) head[b] = t; else nxt[last] = t;
            }
            while (t + 1 >= sc) {                                       // addCount -> transfer (this thread's bins: i = r mod 16)
                for (int i = r; i < cap; i += 16) {
                    const int f = head[i];
                    int ln = -1, hn = -1;
                    if (f >= 0) {
                        unsigned run_bit = hsh[f] & (unsigned)cap;
                        int last_run = f;
                        for (int p = nxt[f]; p >= 0; p = nxt[p]) { const unsigned bb = hsh[p] & (unsigned)cap; if (bb != run_bit) { run_bit = bb; last_run = p; } }
                        if (run_bit == 0) ln = last_run; else hn = last_run;
                        for (int p = f; p != last_run;) {
                            const int pn = nxt[p];
                            if ((hsh[p] & (unsigned)cap) == 0) { nxt[p] = ln; ln = p; } else { nxt[p] = hn; hn = p; }
                            p = pn;
                        }
                    }
                    alt[i] = ln; alt[i + cap] = hn;
                }
                int *tsw = head; head = alt; alt = tsw;
                cap *= 2; sc = cap - (cap >> 2);
            }
        }
        if (lb) s_long = 1;
    }
    __syncthreads();
    {                                                                   // the map's iteration: bins ascending, chains in link order
        __shared__ int s_b[DEEP_THREADS + 1];
        const int per = (F + DEEP_THREADS - 1) / DEEP_THREADS, lo = tid * per, hi = lo + per < F ? lo + per : F;
        int c = 0;
        for (int i = lo; i < hi; i++) for (int p = A[i]; p >= 0; p = nxt[p]) c++;
        s_b[tid] = c;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int t = 0; t < DEEP_THREADS; t++) { const int v = s_b[t]; s_b[t] = acc; acc += v; }
            if (s_long) *long_bin = 1;
        }
        __syncthreads();
        int o = s_b[tid];
        for (int i = lo; i < hi; i++) for (int p = A[i]; p >= 0; p = nxt[p]) out[o++] = keys[p];
    }
    __syncthreads();
    return true;
}
constexpr int PAR_MIN = 192;                   // below this many keys one thread is faster than the CTA's barriers

// ---- keys of clusterLocal in the iteration order of the Int2ObjectOpenHashMap (team leader) ---------------------------------------------------
template <int SM_INTS>
__device__ void seq_keys(const int *idx, int L, DeepW &w, int *smem)     // collective over the team's first CTA
{
    const int nk = compact_par(L, [&](int i) { return w.cnt[idx ? idx[i] : i] > 1; }, [&](int i) { return idx ? idx[i] : i; }, w.tmp);
    if (threadIdx.x == 0) w.hdr[H_NK] = nk;
    const TabBufs tb = bufs_for<SM_INTS>(nk, smem, w);
    if (nk >= PAR_MIN) {
        int hz;
        const int tn = fu_fill_par(w.tmp, nk, tb.A, w.first, w.corder, &hz);
        fu_iter_par(tb.A, tn, hz, w.keys);
    } else if (threadIdx.x == 0) {
        int *tab, hz;
        const int tn = fu_fill_into_a(w.tmp, nk, tb, &tab, &hz);
        fu_iter(tab, tn, hz, w.keys);
    }
    __syncthreads();
}

// ---- groups of the keys by their chosen entry -> clusters appended to the job's list ------------------------------------------------------------
// G1 (parallel over the key positions): entry of every key, first position / size / member sum of every entry
template <int CS>
__device__ void groups_scan(DeepW &w)
{
    const int nk = w.hdr[H_NK];
    for (int p = Team<CS>::tid(); p < nk; p += Team<CS>::size()) {
        const int x = w.keys[p], e = w.chosen[x];
        w.ck[p] = e;
        atomicMin(&w.firstpos[e], p);
        atomicAdd(&w.gcount[e], 1);
        atomicAdd(&w.gsum[e], (unsigned)x);
    }
}
template <int CS>
__device__ void groups_first(DeepW &w)
{
    const int nk = w.hdr[H_NK];
    for (int p = Team<CS>::tid(); p < nk; p += Team<CS>::size()) w.isfirst[p] = w.firstpos[w.ck[p]] == p;
}
// G2 (the team's first CTA; its first thread after the compaction): the groups in first-seen order -> ConcurrentHashMap order -> HashSet<Set<Integer>> order -> cluster slots.  round 1 applies the
// depth rule (L77-L84), round 2 keeps the groups of more than one read (L109)
template <int SM_INTS>
__device__ void groups_layout(int round, int fold_depth, DeepW &w, int *smem)
{
    const int nk = w.hdr[H_NK];
    const int ng = compact_par(nk, [&](int p) { return w.isfirst[p] != 0; }, [&](int p) { return w.ck[p]; }, w.first);
    __shared__ int s_long, s_maxdepth;
    if (threadIdx.x == 0) { s_long = 0; s_maxdepth = 0; }
    __syncthreads();
    int lb = 0;
    if (!chm_order_par<SM_INTS>(w.first, ng, w.corder, w.perm, smem, &lb)) {
        if (threadIdx.x == 0) chm_order(w.first, ng, w.corder, w.chmA, w.chmB, w.chm_nxt, w.chm_h, &lb);
        __syncthreads();
    }
    if (lb) s_long = 1;
    int *gk = w.gm;                                                     // sizes of the groups in map order (gm is free until groups_build)
    for (int t = threadIdx.x; t < ng; t += DEEP_THREADS) {
        const int e = w.corder[t], k = w.gcount[e];
        w.ht[t] = w.gsum[e];
        gk[t] = k;
        atomicMax(&s_maxdepth, k);
    }
    __syncthreads();
    // HashSet<Set<Integer>> order of the groups (hash = member sum): counting sort by bucket, chains in insertion (= map) order
    const int cap = dp_jdk_cap(ng);
    if (2 * (cap + 1) <= SM_INTS) {
        int *cnt = smem, *cur = smem + cap + 1;
        for (int b = threadIdx.x; b <= cap; b += DEEP_THREADS) cnt[b] = 0;
        __syncthreads();
        for (int t = threadIdx.x; t < ng; t += DEEP_THREADS) atomicAdd(&cnt[(dp_spread(w.ht[t]) & (unsigned)(cap - 1)) + 1], 1);
        __syncthreads();
        if (threadIdx.x == 0) for (int b = 0; b < cap; b++) { if (cnt[b + 1] >= 9) s_long = 1; cnt[b + 1] += cnt[b]; }
        __syncthreads();
        for (int b = threadIdx.x; b < cap; b += DEEP_THREADS) cur[b] = cnt[b];
        __syncthreads();
        for (int t = threadIdx.x; t < ng; t += DEEP_THREADS) w.perm[atomicAdd(&cur[dp_spread(w.ht[t]) & (unsigned)(cap - 1)], 1)] = t;
        __syncthreads();
        for (int b = threadIdx.x; b < cap; b += DEEP_THREADS) {          // a bucket's chain keeps insertion order: sort its few entries by t
            const int lo = cnt[b], hi = cnt[b + 1];
            for (int i = lo + 1; i < hi; i++) {
                const int v = w.perm[i];
                int j = i - 1;
                while (j >= lo && w.perm[j] > v) { w.perm[j + 1] = w.perm[j]; j--; }
                w.perm[j + 1] = v;
            }
        }
    } else if (threadIdx.x == 0) { int l2 = 0; jdk_order(w.ht, ng, w.perm, w.tabA, &l2); if (l2) s_long = 1; }
    __syncthreads();
    int *pe = w.gm + ng, *pk = w.pscr;                                   // entry and size of the groups in HashSet order
    for (int t = threadIdx.x; t < ng; t += DEEP_THREADS) { const int g = w.perm[t]; pe[t] = w.corder[g]; pk[t] = gk[g]; }
    __syncthreads();
    // cluster slots in that order: round 1 applies the depth rule, round 2 keeps the groups of more than one read; every thread lays out a
    // contiguous piece after a prefix sum over (clusters kept, reads in them)
    const int maxdepth = s_maxdepth;
    const int n_cl0 = w.hdr[H_NCL], total0 = w.hdr[H_TOTAL];
    __shared__ int s_nc[DEEP_THREADS + 1], s_nr[DEEP_THREADS + 1], s_nbig;
    const int per = (ng + DEEP_THREADS - 1) / DEEP_THREADS, lo = threadIdx.x * per, hi = lo + per < ng ? lo + per : ng;
    auto kept = [&](int k) { return round == 1 ? ((long long)k * fold_depth > maxdepth) : (k > 1); };
    int c = 0, rsum = 0;
    for (int t = lo; t < hi; t++) if (kept(pk[t])) { c++; rsum += pk[t]; }
    s_nc[threadIdx.x] = c; s_nr[threadIdx.x] = rsum;
    if (threadIdx.x == 0) s_nbig = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ac = 0, ar = 0;
        for (int t = 0; t < DEEP_THREADS; t++) { const int vc = s_nc[t], vr = s_nr[t]; s_nc[t] = ac; s_nr[t] = ar; ac += vc; ar += vr; }
        s_nc[DEEP_THREADS] = ac; s_nr[DEEP_THREADS] = ar;
    }
    __syncthreads();
    int q = n_cl0 + s_nc[threadIdx.x], off = total0 + s_nr[threadIdx.x];
    for (int t = lo; t < hi; t++) {
        const int e = pe[t], k = pk[t];
        if (!kept(k)) { w.grp_q[e] = round == 1 ? -2 : -1; continue; }
        w.grp_q[e] = q;
        w.cl_entry[q] = e; w.cl_beg[q] = off; w.cl_len0[q] = k; w.cl_len[q] = k; w.cl_nvict[q] = 0; w.cl_dirty[q] = 1; w.cl_center[q] = -1;
        if (k > SMALL_K) w.big_list[atomicAdd(&s_nbig, 1)] = q;
        q++; off += k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        w.hdr[H_Q0] = n_cl0;
        w.hdr[H_NG] = ng; w.hdr[H_NCL] = n_cl0 + s_nc[DEEP_THREADS]; w.hdr[H_TOTAL] = total0 + s_nr[DEEP_THREADS]; w.hdr[H_NBIG] = s_nbig;
        if (s_long) w.hdr[H_FLAG] |= 2;
    }
}
// members of one small cluster in the iteration order of its HashSet<Integer> (filled in key order); returns 1 when a bin reached 9 entries
__device__ __forceinline__ int small_jdk_order(const int *in, int k, int *out)
{
    const int cap = dp_jdk_cap(k);
    int o = 0, long_bin = 0;
    for (int b = 0; b < cap; b++) {
        int c = 0;
        for (int i = 0; i < k; i++) if ((int)(dp_spread((unsigned)in[i]) & (unsigned)(cap - 1)) == b) { out[o++] = in[i]; c++; }
        if (c >= 9) long_bin = 1;
    }
    return long_bin;
}
// G3 (parallel): member lists, HashSet order, OneUmiCluster (IntOpenHashSet) order of every new cluster; SKIPPED flags of the dropped groups
template <int CS, int SM_INTS>
__device__ void groups_build(slr_umi_assign_rec *rec, DeepW &w, int *smem)
{
    const int nk = w.hdr[H_NK], n_cl = w.hdr[H_NCL], q0 = w.hdr[H_Q0], n_big = w.hdr[H_NBIG];
    const int tid = Team<CS>::tid(), T = Team<CS>::size();
    for (int p = tid; p < nk; p += T) {                                 // flagDontUMIassignRecords (ClusterOneBase.java:L57, L71)
        const int e = w.ck[p];
        if (w.grp_q[e] == -2) {
            const int k = w.gcount[e], x = w.keys[p];
            rec[x].flags |= SLR_UA_SKIPPED; rec[x].cluster_size = (uint16_t)(k > 65535 ? 65535 : k);
        }
    }
    int long_bin = 0;
    const int n_ctas_s = T / DEEP_THREADS;                               // clusters dealt round-robin to the CTAs: every thread scans the key list
    for (int qq = Team<CS>::rank() + n_ctas_s * (int)threadIdx.x; q0 + qq < n_cl; qq += T) {     // small clusters: one thread each, local tables
        const int q = q0 + qq;
        const int k = w.cl_len0[q];
        if (k > SMALL_K) continue;
        const int e = w.cl_entry[q], base = w.cl_beg[q];
        int arr[SMALL_K], ord[SMALL_K], tab[32];
        int c = 0;
        for (int p0 = w.firstpos[e]; c < k; p0 += 8) {                  // 8 loads of the scan in flight (the exit test would serialise them)
            int v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = p0 + u < nk ? w.ck[p0 + u] : -1;
#pragma unroll
            for (int u = 0; u < 8; u++) if (v[u] == e && c < k) arr[c++] = w.keys[p0 + u];
        }
        long_bin |= small_jdk_order(arr, k, ord);
        int *tp, hz;
        const int tn = fu_fill(ord, k, tab, tab, &tp, &hz);
        fu_iter(tp, tn, hz, arr);
        for (int i = 0; i < k; i++) { w.mem[base + i] = ord[i]; w.it[base + i] = arr[i]; w.pos_cl[base + i] = q; w.clid[ord[i]] = q; }
    }
    // big clusters: one CTA each (its first thread walks the tables, which live in shared memory when they fit; clusters whose tables need the
    // job's single global pair all go to CTA 0)
    const int n_ctas = T / DEEP_THREADS, cta = Team<CS>::rank();
    for (int bi = 0; bi < n_big; bi++) {
        const int q = w.big_list[bi], k = w.cl_len0[q];
        const bool sm = fits_smem<SM_INTS>(k);
        if ((sm ? bi % n_ctas : 0) != cta) continue;
        const int e = w.cl_entry[q], base = w.cl_beg[q];
        if (threadIdx.x == 0) {
            int c = 0;
            for (int p0 = w.firstpos[e]; c < k; p0 += 8) {
                int v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = p0 + u < nk ? w.ck[p0 + u] : -1;
#pragma unroll
                for (int u = 0; u < 8; u++) if (v[u] == e && c < k) w.gm[base + c++] = w.keys[p0 + u];
            }
            const TabBufs tb = bufs_for<SM_INTS>(k, smem, w);
            jdk_order(reinterpret_cast<const unsigned *>(w.gm + base), k, w.pscr + base, tb.cnt, &long_bin);
            for (int i = 0; i < k; i++) w.mem[base + i] = w.gm[base + w.pscr[base + i]];
            if (k < PAR_MIN) {
                int *tp, hz;
                const int tn = fu_fill_into_a(w.mem + base, k, tb, &tp, &hz);
                fu_iter(tp, tn, hz, w.it + base);
            }
        }
        __syncthreads();
        if (k >= PAR_MIN) {
            const TabBufs tb = bufs_for<SM_INTS>(k, smem, w);
            int hz;
            const int tn = fu_fill_par(w.mem + base, k, tb.A, w.gm + base, w.sumsq + base, &hz);
            fu_iter_par(tb.A, tn, hz, w.it + base);
        }
        for (int i = threadIdx.x; i < k; i += DEEP_THREADS) { w.pos_cl[base + i] = q; w.clid[w.mem[base + i]] = q; }
        __syncthreads();
    }
    if (long_bin) atomicOr(&w.hdr[H_FLAG], 2);
}

// ---- off-centre removal ----------------------------------------------------------------------------------------------------------------------------
// R1 (the team's first CTA): the unclustered reads in ascending order, then room for the removed ones cluster by cluster (list order)
__device__ void remove_layout(int n, DeepW &w)
{
    const int n_cl = w.hdr[H_NCL];
    const int nu0 = compact_par(n, [&](int d) { return w.clid[d] < 0; }, [&](int d) { return d; }, w.idx);
    // clusters that lose members, in list order (tmp is free here)
    const int n_hit = compact_par(n_cl, [&](int c) { return w.cl_nvict[c] > 0; }, [&](int c) { return c; }, w.tmp);
    for (int c = threadIdx.x; c < n_cl; c += DEEP_THREADS) w.cl_dirty[c] = w.cl_nvict[c] > 0;
    if (threadIdx.x != 0) return;
    int nu = nu0, n_removed = 0, n_big = 0;
    for (int i0 = 0; i0 < n_hit; i0 += 8) {
        int cc[8], vv[8], ll[8];
#pragma unroll
        for (int u = 0; u < 8; u++) cc[u] = i0 + u < n_hit ? w.tmp[i0 + u] : 0;
#pragma unroll
        for (int u = 0; u < 8; u++) { vv[u] = w.cl_nvict[cc[u]]; ll[u] = w.cl_len[cc[u]]; }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (i0 + u >= n_hit) break;
            w.rem_off[cc[u]] = nu; nu += vv[u]; n_removed += vv[u];
            if (ll[u] > SMALL_K) w.big_list[n_big++] = cc[u];
        }
    }
    w.hdr[H_NU] = nu; w.hdr[H_NREM] = n_removed; w.hdr[H_NBIG] = n_big;
}
// R2 (parallel over the clusters that lose members): OneUmiCluster.removeEntries (OneUmiCluster.java:L114-L119)
template <int CS, int SM_INTS>
__device__ void remove_apply(DeepW &w, int *smem)
{
    const int n_cl = w.hdr[H_NCL], n_big = w.hdr[H_NBIG];
    const int tid = Team<CS>::tid(), T = Team<CS>::size();
    for (int c = Team<CS>::rank() + (T / DEEP_THREADS) * (int)threadIdx.x; c < n_cl; c += T) {
        const int k = w.cl_len[c];
        if (w.cl_nvict[c] == 0 || k > SMALL_K) continue;
        const int base = w.cl_beg[c];
        int arr[SMALL_K], tab[32], wrapped[SMALL_K];
        for (int i = 0; i < k; i++) arr[i] = w.mem[base + i];
        int *tp, hz, *alt = tab, tn, size = k, o = w.rem_off[c];
        tn = fu_fill(arr, k, tab, tab, &tp, &hz);
        fu_iter(tp, tn, hz, arr);
        for (int i = 0; i < k; i++) if (w.victim[arr[i]]) { w.idx[o++] = arr[i]; w.clid[arr[i]] = -1; }
        fu_remove_all(&tp, &alt, &tn, &size, &hz, w.victim, wrapped);
        const int k2 = fu_iter(tp, tn, hz, w.it + base);
        for (int i = 0; i < k; i++) w.victim[arr[i]] = 0;
        w.cl_len[c] = k2; w.cl_nvict[c] = 0;
    }
    const int n_ctas = T / DEEP_THREADS, cta = Team<CS>::rank();
    for (int bi = 0; bi < n_big; bi++) {
        const int c = w.big_list[bi], k = w.cl_len[c];
        const bool sm = fits_smem<SM_INTS>(k);
        if ((sm ? bi % n_ctas : 0) != cta) continue;
        const int base = w.cl_beg[c];
        const TabBufs tb = bufs_for<SM_INTS>(k, smem, w);
        __shared__ int s_tn, s_hz;
        if (k >= PAR_MIN) {
            int hz;
            const int tn = fu_fill_par(w.mem + base, k, tb.A, w.gm + base, w.sumsq + base, &hz);
            if (threadIdx.x == 0) { s_tn = tn; s_hz = hz; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int o = w.rem_off[c];
            for (int i = 0; i < k; i++) { const int x = w.it[base + i]; if (w.victim[x]) { w.idx[o++] = x; w.clid[x] = -1; } }
            int *tp = tb.A, hz = s_hz, tn = s_tn, size = k;
            if (k < PAR_MIN) tn = fu_fill_into_a(w.mem + base, k, tb, &tp, &hz);
            int *alt = tp == tb.A ? tb.B : tb.A;
            fu_remove_all(&tp, &alt, &tn, &size, &hz, w.victim, w.pscr + base);
            w.cl_len[c] = fu_iter(tp, tn, hz, w.it + base);
            w.cl_nvict[c] = 0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += DEEP_THREADS) w.victim[w.mem[base + i]] = 0;
        __syncthreads();
    }
}

template <int CS, int SM_INTS>
__device__ void cluster_local_round(const int32_t *__restrict__ M, int n, int round, const slr_umi_assign_params &P, int qv01,
                                    slr_umi_assign_rec *__restrict__ rec, DeepW &w, int *smem)
{
    const bool leader = Team<CS>::tid() == 0;
    const int tid = Team<CS>::tid(), T = Team<CS>::size();
    const int *idx = round == 1 ? nullptr : w.idx, *inU = round == 1 ? nullptr : w.inU;
    const int L = round == 1 ? n : w.hdr[H_NU];
    const int sb = round == 1 ? 0 : 8;
    DEEP_STAMP(w, sb + 0);
    pass_counts<CS>(M, n, P.ed_complete, idx, L, inU, w);
    for (int i = tid; i < n; i += T) { w.firstpos[i] = 0x7FFFFFFF; w.gcount[i] = 0; w.gsum[i] = 0; w.grp_q[i] = -1; }
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 1);
    if (Team<CS>::rank() == 0) seq_keys<SM_INTS>(idx, L, w, smem);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 2);
    if (w.hdr[H_NK] == 0) return;
    unsigned long long *best64 = reinterpret_cast<unsigned long long *>(w.gm);      // 2 (n + 2) words, free until groups_build
    pass_choose_prepare<CS>(w, best64);
    Team<CS>::sync();
    pass_choose_sweep<CS, false>(M, n, P.ed_complete, inU, w, best64);
    Team<CS>::sync();
    pass_choose_sweep<CS, true>(M, n, P.ed_complete, inU, w, best64);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 3);
    groups_scan<CS>(w);
    Team<CS>::sync();
    groups_first<CS>(w);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 4);
    if (Team<CS>::rank() == 0) groups_layout<SM_INTS>(round, P.fold_depth, w, smem);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 5);
    groups_build<CS, SM_INTS>(rec, w, smem);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 6);
    pass_sumsq<CS>(M, n, w);
    Team<CS>::sync();
    pass_centers<CS>(qv01, w);
    Team<CS>::sync();
    DEEP_STAMP(w, sb + 7);
}

template <int CS, int SM_INTS>
__device__ void deep_job(const int32_t *__restrict__ M, int n, const slr_umi_assign_params P, int qv01, slr_umi_assign_rec *__restrict__ rec, int *W,
                         int *smem)
{
    DeepW w;
    carve(W, n, w);
    const int ed = P.ed_complete;
    const bool leader = Team<CS>::tid() == 0;
    const int tid = Team<CS>::tid(), T = Team<CS>::size();
    const int lane = threadIdx.x & 31, warp = tid >> 5, n_warps = T >> 5;
    for (int i = tid; i < n; i += T) { w.clid[i] = -1; w.victim[i] = 0; w.cnt[i] = 0; w.chosen[i] = -1; }
    if (leader) for (int i = 0; i < H_WORDS; i++) w.hdr[i] = 0;
    Team<CS>::sync();
    cluster_local_round<CS, SM_INTS>(M, n, 1, P, qv01, rec, w, smem);     // clusterLocal over all reads, depth rule, centres
    if (w.hdr[H_NCL] > 0) {
        pass_victims<CS>(M, n, ed, w);
        Team<CS>::sync();
        if (Team<CS>::rank() == 0) remove_layout(n, w);
        Team<CS>::sync();
        if (w.hdr[H_NREM] > 0) {
            remove_apply<CS, SM_INTS>(w, smem);
            Team<CS>::sync();
            pass_sumsq<CS>(M, n, w);
            Team<CS>::sync();
            pass_centers<CS>(qv01, w);
            // clusterLocal over the unclustered reads
            const int nu = w.hdr[H_NU];
            for (int i = tid; i < n; i += T) { w.inU[i] = 0; w.cnt[i] = 0; }
            Team<CS>::sync();
            for (int i = tid; i < nu; i += T) w.inU[w.idx[i]] = 1;
            const int ncl = w.hdr[H_NCL];
            for (int c = tid; c < ncl; c += T) w.cl_dirty[c] = 0;
            Team<CS>::sync();
            cluster_local_round<CS, SM_INTS>(M, n, 2, P, qv01, rec, w, smem);
        }
        Team<CS>::sync();
        // ---- per cluster: mean shift against the centre and the number of members within ED of it (L126-L139)
        DEEP_STAMP(w, 14);
        const int n_cl = w.hdr[H_NCL];
        for (int c = warp; c < n_cl; c += n_warps) {
            const int m = w.cl_len[c], b = w.cl_beg[c], center = w.cl_center[c];
            int sum = 0, nf = 0;
            if (m > 1)
                for (int q = lane; q < m; q += 32) {
                    const int x = w.it[b + q];
                    if (x != center) sum += dp_pos1_offset(M[(size_t)center * n + x]);
                    if (dp_ed(M[(size_t)x * n + center]) <= ed) nf++;
                }
            sum = __reduce_add_sync(FULLM, sum); nf = __reduce_add_sync(FULLM, nf);
            if (lane == 0) {
                w.cl_nf[c] = nf;
                w.cl_offmean[c] = m > 1 ? (int)floor((double)sum / (double)(m - 1) + 0.5) : 0;
            }
        }
        Team<CS>::sync();
        // ---- per read (L145-L162)
        const int total = w.hdr[H_TOTAL];
        for (int pos = warp; pos < total; pos += n_warps) {
            const int c = w.pos_cl[pos];
            const int m = w.cl_len[c];
            if (pos - w.cl_beg[c] >= m || m <= 1 || w.cl_nf[c] <= 1) continue;
            const int x = w.it[pos], center = w.cl_center[c];
            if (dp_ed(M[(size_t)x * n + center]) > ed) continue;
            if (rec[x].flags & SLR_UA_SKIPPED) continue;                  // ClusterOneBase.java:L122-L123
            int best = 127;
            if (n_cl > 1) {
                const int32_t *row = M + (size_t)x * n;
                for (int y0 = 0; y0 < n; y0 += 32 * DU) {
                    int32_t v[DU];
                    int cl[DU];
#pragma unroll
                    for (int u = 0; u < DU; u++) {
                        const int y = y0 + u * 32 + lane;
                        v[u] = y < n ? __ldg(row + y) : 0x7F;
                        cl[u] = y < n ? w.clid[y] : c;
                    }
#pragma unroll
                    for (int u = 0; u < DU; u++)
                        if (cl[u] != c) { const int e = dp_ed(v[u]); best = e < best ? e : best; }
                }
                best = __reduce_min_sync(FULLM, best);
            }
            if (lane == 0) {
                const int32_t cell = M[(size_t)center * n + x];
                slr_umi_assign_rec r = rec[x];
                r.center = center; r.u1 = (int8_t)dp_ed(cell); r.u2 = (int8_t)((n_cl > 1 && best != 127) ? best : -1); r.pos2 = (int8_t)dp_pos2_code(cell);
                r.offset_center_mean = (int8_t)w.cl_offmean[c]; r.flags |= SLR_UA_ASSIGNED; r.cluster_size = (uint16_t)(m > 65535 ? 65535 : m);
                rec[x] = r;
            }
        }
    }
    Team<CS>::sync();
    const int n_cl = w.hdr[H_NCL], flag = w.hdr[H_FLAG];
    for (int i = tid; i < n; i += T) { rec[i].n_clusters = n_cl; if (flag) rec[i].flags |= SLR_UA_TIE_UNPIN; }
    Team<CS>::sync();
    DEEP_STAMP(w, 15);
}

template <int CS, int SM_INTS>
__global__ void __launch_bounds__(DEEP_THREADS) umi_assign_deep_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                       const long long *__restrict__ ooff, const slr_umi_assign_params P,
                                                                       const uint8_t *__restrict__ job_qv01, slr_umi_assign_rec *__restrict__ rec,
                                                                       const int32_t *__restrict__ list, const long long *__restrict__ list_off,
                                                                       const unsigned int *__restrict__ count, int *__restrict__ scratch)
{
    extern __shared__ int deep_smem[];
    const unsigned int total = *count;
    const unsigned int team = CS == 0 ? 0 : blockIdx.x / (CS == 0 ? 1 : CS), n_teams = CS == 0 ? 1 : gridDim.x / (CS == 0 ? 1 : CS);
    for (unsigned int k = team; k < total; k += n_teams) {
        const long long j = list[k];
        const long long r0 = joff[j];
        const int n = (int)(joff[j + 1] - r0);
        if ((CS == 8 && n > SLR_UA_DEEP_MEDIUM) || (CS == 0 && n <= SLR_UA_DEEP_MEDIUM)) continue;      // the two big classes share one list
        deep_job<CS, SM_INTS>(mat + ooff[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0, scratch + list_off[k], deep_smem);
    }
}

constexpr int SM_SMALL = 3 * 1024, SM_BIG = 48 * 1024;     // words of dynamic shared memory of the one-CTA / multi-CTA teams

}  // namespace

// lists: the deep jobs filed by umi_assign_init (class 2: up to SLR_UA_DEEP_SMALL reads, one CTA each; class 3: a cluster of 8 CTAs each)
cudaError_t slr_launch_umi_assign_deep(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets,
                                       const slr_umi_assign_params &P, const uint8_t *d_job_qv01, slr_umi_assign_rec *d_rec,
                                       const int32_t *d_list_small, const long long *d_off_small, const unsigned int *d_count_small,
                                       const int32_t *d_list_big, const long long *d_off_big, const unsigned int *d_count_big, int *d_words,
                                       long long max_jobs, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // function attributes belong to the device (one module per context): set once per device, not once per process
    static std::mutex attr_mtx;
    static bool attr_done[64];
    {
        std::lock_guard<std::mutex> lk(attr_mtx);
        if (!attr_done[dev & 63]) {
            cudaError_t ae = cudaFuncSetAttribute(umi_assign_deep_kernel<8, SM_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BIG * 4);
            if (ae == cudaSuccess)
                ae = cudaFuncSetAttribute(umi_assign_deep_kernel<0, SM_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BIG * 4);
            if (ae != cudaSuccess) return ae;
            attr_done[dev & 63] = true;
        }
    }
    long long gs = max_jobs < (long long)sms * 2 ? max_jobs : (long long)sms * 2;
    if (gs < 1) gs = 1;
    umi_assign_deep_kernel<1, SM_SMALL><<<(unsigned)gs, DEEP_THREADS, SM_SMALL * 4, stream>>>(d_mat, d_job_offsets, d_out_offsets, P, d_job_qv01, d_rec,
                                                                                           d_list_small, d_off_small, d_count_small, d_words);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    long long teams = max_jobs < (long long)sms / 8 ? max_jobs : (long long)sms / 8;
    if (teams < 1) teams = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(teams * 8)); cfg.blockDim = dim3(DEEP_THREADS); cfg.dynamicSmemBytes = SM_BIG * 4; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, umi_assign_deep_kernel<8, SM_BIG>, d_mat, d_job_offsets, d_out_offsets, P, d_job_qv01, d_rec, d_list_big, d_off_big,
                           d_count_big, d_words);
    if (e != cudaSuccess) return e;
    // giant jobs (above SLR_UA_DEEP_MEDIUM reads): one after the other on the whole GPU, grid-wide barriers between the phases
    void *args[] = {(void *)&d_mat, (void *)&d_job_offsets, (void *)&d_out_offsets, (void *)&P, (void *)&d_job_qv01, (void *)&d_rec,
                    (void *)&d_list_big, (void *)&d_off_big, (void *)&d_count_big, (void *)&d_words};
    return cudaLaunchCooperativeKernel((const void *)umi_assign_deep_kernel<0, SM_BIG>, dim3((unsigned)sms), dim3(DEEP_THREADS), args, SM_BIG * 4, stream);
}
