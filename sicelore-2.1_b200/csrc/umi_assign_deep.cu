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
// One TEAM per job: one CTA (jobs up to DEEP_SMALL reads) or a thread-block cluster of 8 CTAs.  The O(n^2) passes over the matrix (neighbour
// counts, entry choice, sums of squared distances, U2) are spread over the team's warps with coalesced row reads; the hash-table emulations are
// inherently sequential and run on the team's first thread between team barriers.  All working arrays live in the caller's scratch.
#include <cooperative_groups.h>
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
enum { H_NK = 0, H_NG, H_NCL, H_NU, H_NREM, H_FLAG, H_TOTAL, H_ROUND2_FIRST, H_NDIRTY, H_WORDS = 16 };

struct DeepW {
    int *hdr, *cnt, *chosen, *keys, *tmp, *gid, *first, *gsz, *gm, *fill, *corder, *perm, *mem, *it, *pos_cl, *sumsq;
    int *cl_beg, *cl_len0, *cl_len, *cl_center, *cl_nvict, *cl_dirty, *cl_offmean, *cl_nf;
    int *clid, *victim, *idx, *inU, *tabA, *tabB, *wrapped, *chmA, *chmB, *chm_nxt;
    unsigned *hs, *chm_h;
    unsigned long long *sig;
};
__device__ inline void carve(int *W, int n, DeepW &w)
{
    const int s = n + 2;
    w.sig = reinterpret_cast<unsigned long long *>(W);             // 8-byte aligned: every arena offset is even
    int *p = W + 2 * s;
    w.hdr = p; p += 64;
    auto take = [&](int words) { int *q = p; p += words; return q; };
    w.cnt = take(s); w.chosen = take(s); w.keys = take(s); w.tmp = take(s); w.gid = take(s); w.first = take(s); w.gsz = take(s); w.gm = take(s);
    w.fill = take(s); w.corder = take(s); w.perm = take(s);
    w.mem = take(2 * s); w.it = take(2 * s); w.pos_cl = take(2 * s); w.sumsq = take(2 * s);
    w.cl_beg = take(s); w.cl_len0 = take(s); w.cl_len = take(s); w.cl_center = take(s); w.cl_nvict = take(s); w.cl_dirty = take(s);
    w.cl_offmean = take(s); w.cl_nf = take(s);
    w.clid = take(s); w.victim = take(s); w.idx = take(2 * s); w.inU = take(s); w.wrapped = take(s); w.chm_nxt = take(s);
    w.hs = reinterpret_cast<unsigned *>(take(s)); w.chm_h = reinterpret_cast<unsigned *>(take(s));
    // 11 + 8 + 8 + 7 + 2 = 36 single + 4 double + idx double = 36 + 8 + 2 = 46 s  (<= 50 s)
    w.tabA = take(3 * n + 64); w.tabB = take(3 * n + 64); w.chmA = take(4 * n + 64); w.chmB = take(4 * n + 64);
}

// ---- team = 1 CTA or a cluster of CS CTAs ---------------------------------------------------------------------------------------------------
template <int CS>
struct Team {
    __device__ static __forceinline__ void sync()
    {
        if (CS == 1) __syncthreads();
        else { __threadfence(); cg::this_cluster().sync(); }
    }
    __device__ static __forceinline__ int rank() { return CS == 1 ? 0 : (int)cg::this_cluster().block_rank(); }
    __device__ static __forceinline__ int tid() { return rank() * DEEP_THREADS + (int)threadIdx.x; }
    __device__ static __forceinline__ int size() { return CS * DEEP_THREADS; }
};

// ---- sequential container emulations (one thread) ---------------------------------------------------------------------------------------------
// fastutil open-addressing table filled in the given order (keys distinct); returns the table size, *tab_out = the table, *has_zero
__device__ int fu_fill(const int *in, int k, int *bufA, int *bufB, int **tab_out, int *has_zero)
{
    int n = 32, size = 0, hz = 0;
    int *tab = bufA, *alt = bufB;
    for (int i = 0; i < n; i++) tab[i] = 0;
    for (int i = 0; i < k; i++) {
        const int key = in[i];
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
template <int CS>
__device__ void pass_counts(const int32_t *__restrict__ M, int n, int ed, const int *idx, int L, const int *inU, DeepW &w)
{
    const int lane = threadIdx.x & 31, warp = Team<CS>::tid() >> 5, n_warps = Team<CS>::size() >> 5;
    for (int i = warp; i < L; i += n_warps) {
        const int a = idx ? idx[i] : i;
        const int32_t *row = M + (size_t)a * n;
        int c = 0;
        unsigned long long s = 0;
        for (int j = lane; j < n; j += 32)
            if ((!inU || inU[j]) && dp_ed(row[j]) <= ed) { c++; s += dp_sig(j); }
        c = __reduce_add_sync(FULLM, c);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULLM, s, o);
        if (lane == 0) { w.cnt[a] = c; w.sig[a] = s; }
    }
}
// every key c joins the first entry (in key order) with the largest neighbour set that contains c
template <int CS>
__device__ void pass_choose(const int32_t *__restrict__ M, int n, int ed, const int *inU, DeepW &w)
{
    const int nk = w.hdr[H_NK];
    int harmful = 0;
    for (int c = Team<CS>::tid(); c < n; c += Team<CS>::size()) {
        if ((inU && !inU[c]) || w.cnt[c] <= 1) continue;
        int best = -1, bestcnt = -1, tie_diff = 0;
        unsigned long long bestsig = 0;
        for (int p = 0; p < nk; p++) {
            const int e = w.keys[p];
            if (dp_ed(M[(size_t)e * n + c]) > ed) continue;
            const int ce = w.cnt[e];
            if (ce > bestcnt) { best = e; bestcnt = ce; bestsig = w.sig[e]; tie_diff = 0; }
            else if (ce == bestcnt && w.sig[e] != bestsig) tie_diff = 1;
        }
        w.chosen[c] = best;
        harmful |= tie_diff;
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
            for (int j = lane; j < n; j += 32)
                if (w.clid[j] == c && j != a) { const int e = dp_ed(row[j]); s += e * e; }
        } else {
            const int *seg = w.it + w.cl_beg[c];
            for (int q = lane; q < m; q += 32) { const int x = seg[q]; if (x != a) { const int e = dp_ed(row[x]); s += e * e; } }
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

// ---- sequential phases -------------------------------------------------------------------------------------------------------------------------
__device__ void seq_keys(int n, const int *idx, int L, DeepW &w)
{
    int nk = 0;
    for (int i = 0; i < L; i++) { const int a = idx ? idx[i] : i; if (w.cnt[a] > 1) w.tmp[nk++] = a; }
    int *tab, hz;
    const int tn = fu_fill(w.tmp, nk, w.tabA, w.tabB, &tab, &hz);      // Int2ObjectOpenHashMap: the set's layout and iteration
    fu_iter(tab, tn, hz, w.keys);
    w.hdr[H_NK] = nk;
}
// groups of the keys by their chosen entry -> clusters appended to the job's list.  round 1 applies the depth rule, round 2 keeps sizes > 1
__device__ void seq_groups(int n, int round, int fold_depth, slr_umi_assign_rec *rec, DeepW &w)
{
    const int nk = w.hdr[H_NK];
    int ng = 0, long_bin = 0;
    for (int i = 0; i < n; i++) w.gid[i] = -1;
    for (int i = 0; i < nk; i++) { const int e = w.chosen[w.keys[i]]; if (w.gid[e] < 0) { w.gid[e] = ng; w.first[ng++] = e; } }
    for (int g = 0; g <= ng; g++) w.gsz[g] = 0;
    for (int i = 0; i < nk; i++) w.gsz[w.gid[w.chosen[w.keys[i]]] + 1]++;
    for (int g = 0; g < ng; g++) w.gsz[g + 1] += w.gsz[g];
    for (int g = 0; g < ng; g++) w.fill[g] = w.gsz[g];
    for (int i = 0; i < nk; i++) w.gm[w.fill[w.gid[w.chosen[w.keys[i]]]]++] = w.keys[i];
    chm_order(w.first, ng, w.corder, w.chmA, w.chmB, w.chm_nxt, w.chm_h, &long_bin);
    int maxdepth = 0;
    for (int t = 0; t < ng; t++) {
        const int g = w.gid[w.corder[t]];
        unsigned h = 0;
        for (int i = w.gsz[g]; i < w.gsz[g + 1]; i++) h += (unsigned)w.gm[i];
        w.hs[t] = h;
        if (w.gsz[g + 1] - w.gsz[g] > maxdepth) maxdepth = w.gsz[g + 1] - w.gsz[g];
    }
    jdk_order(w.hs, ng, w.perm, w.tabA, &long_bin);
    int n_cl = w.hdr[H_NCL], total = w.hdr[H_TOTAL];
    for (int t = 0; t < ng; t++) {
        const int g = w.gid[w.corder[w.perm[t]]], k = w.gsz[g + 1] - w.gsz[g];
        const int *src = w.gm + w.gsz[g];
        if (round == 2 && k <= 1) continue;
        // HashSet<Integer> of the group, filled in key order: hash = value
        unsigned *hh = reinterpret_cast<unsigned *>(w.tmp);
        for (int i = 0; i < k; i++) hh[i] = (unsigned)src[i];
        int *pp = w.fill;                                               // free by now (k <= nk <= n)
        jdk_order(hh, k, pp, w.tabA, &long_bin);
        if (round == 1 && !((long long)k * fold_depth > maxdepth)) {
            for (int i = 0; i < k; i++) { rec[src[i]].flags |= SLR_UA_SKIPPED; rec[src[i]].cluster_size = (uint16_t)(k > 65535 ? 65535 : k); }
            continue;
        }
        int *mem = w.mem + total;
        for (int i = 0; i < k; i++) mem[i] = src[pp[i]];
        int *tab, hz;
        const int tn = fu_fill(mem, k, w.tabA, w.tabB, &tab, &hz);      // toCollection(OneUmiCluster::new)
        fu_iter(tab, tn, hz, w.it + total);
        for (int i = 0; i < k; i++) { w.clid[mem[i]] = n_cl; w.pos_cl[total + i] = n_cl; }
        w.cl_beg[n_cl] = total; w.cl_len0[n_cl] = k; w.cl_len[n_cl] = k; w.cl_nvict[n_cl] = 0; w.cl_dirty[n_cl] = 1; w.cl_center[n_cl] = -1;
        n_cl++; total += k;
    }
    w.hdr[H_NG] = ng; w.hdr[H_NCL] = n_cl; w.hdr[H_TOTAL] = total;
    if (long_bin) w.hdr[H_FLAG] |= 2;
}
// unclustered list + off-centre removal (the clusters are visited in list order)
__device__ void seq_remove(int n, DeepW &w)
{
    int nu = 0, n_removed = 0;
    for (int d = 0; d < n; d++) if (w.clid[d] < 0) w.idx[nu++] = d;
    const int n_cl = w.hdr[H_NCL];
    for (int c = 0; c < n_cl; c++) {
        w.cl_dirty[c] = 0;
        if (w.cl_nvict[c] == 0) continue;
        const int b = w.cl_beg[c], k = w.cl_len[c];
        for (int q = 0; q < k; q++) { const int x = w.it[b + q]; if (w.victim[x]) { w.idx[nu++] = x; w.clid[x] = -1; n_removed++; } }
        int *tab, hz, *alt;
        int tn = fu_fill(w.mem + b, k, w.tabA, w.tabB, &tab, &hz), size = k;
        alt = tab == w.tabA ? w.tabB : w.tabA;
        fu_remove_all(&tab, &alt, &tn, &size, &hz, w.victim, w.wrapped);
        const int k2 = fu_iter(tab, tn, hz, w.it + b);
        for (int q = 0; q < k; q++) w.victim[w.mem[b + q]] = 0;
        w.cl_len[c] = k2; w.cl_dirty[c] = 1; w.cl_nvict[c] = 0;
    }
    w.hdr[H_NU] = nu; w.hdr[H_NREM] = n_removed;
}

template <int CS>
__device__ void deep_job(const int32_t *__restrict__ M, int n, const slr_umi_assign_params P, int qv01, slr_umi_assign_rec *__restrict__ rec, int *W)
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
    // ---- round 1: clusterLocal over all reads
    pass_counts<CS>(M, n, ed, nullptr, n, nullptr, w);
    Team<CS>::sync();
    if (leader) seq_keys(n, nullptr, n, w);
    Team<CS>::sync();
    if (w.hdr[H_NK] > 0) {
        pass_choose<CS>(M, n, ed, nullptr, w);
        Team<CS>::sync();
        if (leader) seq_groups(n, 1, P.fold_depth, rec, w);
        Team<CS>::sync();
        pass_sumsq<CS>(M, n, w);
        Team<CS>::sync();
        pass_centers<CS>(qv01, w);
        Team<CS>::sync();
        pass_victims<CS>(M, n, ed, w);
        Team<CS>::sync();
        if (leader) seq_remove(n, w);
        Team<CS>::sync();
        if (w.hdr[H_NREM] > 0) {
            pass_sumsq<CS>(M, n, w);
            Team<CS>::sync();
            pass_centers<CS>(qv01, w);
            // ---- round 2: clusterLocal over the unclustered reads
            const int nu = w.hdr[H_NU];
            for (int i = tid; i < n; i += T) { w.inU[i] = 0; w.cnt[i] = 0; }
            Team<CS>::sync();
            for (int i = tid; i < nu; i += T) w.inU[w.idx[i]] = 1;
            if (leader) { const int ncl = w.hdr[H_NCL]; for (int c = 0; c < ncl; c++) w.cl_dirty[c] = 0; w.hdr[H_ROUND2_FIRST] = ncl; }
            Team<CS>::sync();
            pass_counts<CS>(M, n, ed, w.idx, nu, w.inU, w);
            Team<CS>::sync();
            if (leader) seq_keys(n, w.idx, nu, w);
            Team<CS>::sync();
            if (w.hdr[H_NK] > 0) {
                pass_choose<CS>(M, n, ed, w.inU, w);
                Team<CS>::sync();
                if (leader) seq_groups(n, 2, P.fold_depth, rec, w);
                Team<CS>::sync();
                pass_sumsq<CS>(M, n, w);
                Team<CS>::sync();
                pass_centers<CS>(qv01, w);
            }
        }
        Team<CS>::sync();
        // ---- per cluster: mean shift against the centre and the number of members within ED of it (L126-L139)
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
                for (int y = lane; y < n; y += 32) if (w.clid[y] != c) { const int e = dp_ed(row[y]); best = e < best ? e : best; }
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
}

template <int CS>
__global__ void __launch_bounds__(DEEP_THREADS) umi_assign_deep_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                       const long long *__restrict__ ooff, const slr_umi_assign_params P,
                                                                       const uint8_t *__restrict__ job_qv01, slr_umi_assign_rec *__restrict__ rec,
                                                                       const int32_t *__restrict__ list, const long long *__restrict__ list_off,
                                                                       const unsigned int *__restrict__ count, int *__restrict__ scratch)
{
    const unsigned int total = *count;
    const unsigned int team = blockIdx.x / CS, n_teams = gridDim.x / CS;
    for (unsigned int k = team; k < total; k += n_teams) {
        const long long j = list[k];
        const long long r0 = joff[j];
        const int n = (int)(joff[j + 1] - r0);
        deep_job<CS>(mat + ooff[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0, scratch + list_off[k]);
    }
}

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
    long long gs = max_jobs < (long long)sms * 2 ? max_jobs : (long long)sms * 2;
    if (gs < 1) gs = 1;
    umi_assign_deep_kernel<1><<<(unsigned)gs, DEEP_THREADS, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, P, d_job_qv01, d_rec, d_list_small,
                                                                         d_off_small, d_count_small, d_words);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    long long teams = max_jobs < (long long)sms / 8 ? max_jobs : (long long)sms / 8;
    if (teams < 1) teams = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(teams * 8)); cfg.blockDim = dim3(DEEP_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, umi_assign_deep_kernel<8>, d_mat, d_job_offsets, d_out_offsets, P, d_job_qv01, d_rec, d_list_big, d_off_big,
                              d_count_big, d_words);
}
