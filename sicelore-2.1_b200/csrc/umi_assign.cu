// umi_assign.cu — clustering of the small (cell, region) jobs and the per-read UMI assignment, on the packed matrices in HBM.
//
// Replaces ClusterOneHierarchical.call (F!com/rw/umifinder/analyzers/clustering/ClusterOneHierarchical.class,
// ClusterOneHierarchical.java:L61-L217), the clusterer UmiClustering$Submitter picks for every job of at most 100 reads
// (UmiClustering.java:L239-L261; CLUSTERHOW is the constant DECIDEONCOMPLEXITY, so the pre-grouping branch is dead), for ALL such jobs of a
// BAM chunk at once:
//   reads with a neighbour           DistanceMatrix.generateIndicesWithNeighbours (DistanceMatrix.java:L87-L90)
//   LingPipe complete link           A!com/aliasi/cluster/CompleteLinkClusterer.class (CompleteLinkClusterer.java:L146-L237): every pair in a
//                                    BoundedPriorityQueue (least cost first, among equal costs the pair offered LAST: BoundedPriorityQueue.java
//                                    :L458-L464), merged pair by pair with cost(12, 3) = max(cost(1, 3), cost(2, 3))
//   (single link above complexity_threshold_for_switch_to_single_link_clustering: SingleLinkClusterer.java:L198-L268)
//   Dendrogram.partitionDistance(ed) A!…/Dendrogram.class (Dendrogram.java:L205-L215), clusters of more than one read, depth rule (L121-L127)
//   OneUmiCluster.setClusterCenter   F!com/rw/clustering/OneUmiCluster.class (OneUmiCluster.java:L49-L65): two reads -> by the mean quality of the
//                                    job's first two reads, else least sum of squared distances, first in the set's iteration order
//   per read                         ClusterOneBase.setSamflagsAndStatsForClustered (ClusterOneBase.java:L118-L168): centre, U1 = distance to the
//                                    centre, U2 = least distance to a read outside the cluster, the +-1 shift of matrix[centre][read]; and the
//                                    mean shift of the cluster (ClusterOneHierarchical.java:L143-L147) the caller cuts the centre's UMI with.
// The orders the result depends on are reproduced: the queue's tie rule, java.util.HashSet<Integer> iteration (bucket = value & (capacity - 1),
// chains in insertion order) for memberSet() / transformIndices, and fastutil's IntOpenHashSet (OneUmiCluster) slot order.  The one order the
// JVM itself does not fix — ObjectToSet's HashSet<PairScore> is keyed by identity hash codes — is taken as creation order, and a job whose
// records can depend on it is flagged SLR_UA_TIE_UNPIN (see oracle/slr_oracle_assign.c, the CPU restatement these kernels are tested against).
//
// One warp per job.  Small jobs (<= 32 reads, almost all of them: jobs average 4 reads) run 8 per CTA with 3.6 KB of shared memory each; jobs
// of 33 ... 100 reads run one warp per CTA.  The packed matrix of a job is read once (coalesced), everything else happens in shared memory.
#include "slr_kernels.h"

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ int ua_jdk_cap(int size)
{
    int cap = 16;
    while (size > cap * 3 / 4) cap <<= 1;
    return cap;
}
__device__ __forceinline__ uint32_t ua_fu_mix(int k)
{
    const uint32_t h = (uint32_t)k * 0x9E3779B9u;
    return h ^ (h >> 16);
}
__device__ __forceinline__ int ua_pos1_offset(int32_t p) { return (p & 0x08000000) ? -1 : ((p & 0x10000000) ? 0 : ((p & 0x20000000) ? 1 : 0)); }
__device__ __forceinline__ int ua_pos2_code(int32_t p) { return (p & 0x01000000) ? 0 : ((p & 0x02000000) ? 1 : ((p & 0x04000000) ? 2 : 1)); }

// per-job scratch in shared memory; MAXN = largest job (32 or 100), S = row stride
template <int MAXN>
struct UaShared {
    static constexpr int S = (MAXN + 3) & ~3;
    uint8_t E[MAXN][S];            // ED of every read pair of the job (low byte of the packed cell)
    uint8_t D[MAXN][S];            // complete-link cost between the cluster slots (reduced index space)
    uint8_t R[MAXN][S];            // rank of a pair inside its generation (pairs created by a merge)
    uint8_t iwn[MAXN];             // reduced index -> read
    uint8_t birth[MAXN];           // merge step that created the cluster in a slot (0 = leaf)
    uint8_t head[MAXN], tail[MAXN], nxt[MAXN], csize[MAXN];   // leaf order of a cluster = LinkDendrogram.addMembers order
    uint8_t clid[MAXN];            // read -> cluster ordinal, 0xFF = none
    uint8_t it[MAXN];              // iteration orders of all OneUmiClusters, back to back
    uint8_t tmp[2][MAXN];
    uint8_t tab[MAXN >= 64 ? 256 : 64];   // fastutil table of the cluster a lane is ordering (clusters are ordered one after the other)
    uint8_t cl_start[MAXN / 2 + 1], cl_len[MAXN / 2 + 1], cl_center[MAXN / 2 + 1];
    int8_t cl_off[MAXN / 2 + 1];
    uint32_t idkey[MAXN];
};

// iteration order of a java.util.HashSet<Integer> filled in the order in[0..k): stable by bucket.  Returns true when two share a bucket.
__device__ __noinline__ bool ua_jdk_order(const uint8_t *in, int k, uint8_t *out)
{
    const int cap = ua_jdk_cap(k);
    bool collide = false;
    int o = 0;
    if (cap > 128) {                                         // k > 96 reads with indices below 128 < cap: already bucket-ordered iff ascending
        for (int b = 0; b < 128 && o < k; b++)
            for (int i = 0; i < k; i++)
                if (in[i] == b) out[o++] = in[i];
        return false;
    }
    for (int b = 0; b < cap && o < k; b++) {
        int cnt = 0;
        for (int i = 0; i < k; i++)
            if ((in[i] & (cap - 1)) == b) { out[o++] = in[i]; cnt++; }
        collide |= cnt > 1;
    }
    return collide;
}

// iteration order of a fastutil IntOpenHashSet (default constructor) filled by add() in the order in[0..k): key 0 first, then the slots from
// the last to the first.  tab: 256 bytes (64 when k <= 48).  Keys are read indices < 128.  Up to 24 keys the table keeps its 32 slots and
// its occupancy is one register (no clearing, no rehash): that is every cluster of almost every job.
__device__ __noinline__ void ua_fastutil_order_big(const uint8_t *in, int k, uint8_t *out, uint8_t *tab);
__device__ __forceinline__ void ua_fastutil_order(const uint8_t *in, int k, uint8_t *out, uint8_t *tab)
{
    if (k <= 24) {
        uint32_t occ = 0;
        bool has_zero = false;
        for (int i = 0; i < k; i++) {
            const int key = in[i];
            if (key == 0) { has_zero = true; continue; }
            int pos = (int)(ua_fu_mix(key) & 31u);
            while ((occ >> pos) & 1u) pos = (pos + 1) & 31;
            occ |= 1u << pos;
            tab[pos] = (uint8_t)key;
        }
        int o = 0;
        if (has_zero) out[o++] = 0;
        while (occ) {                                        // slots from the last to the first
            const int pos = 31 - __clz((int)occ);
            occ &= ~(1u << pos);
            out[o++] = tab[pos];
        }
        return;
    }
    ua_fastutil_order_big(in, k, out, tab);
}
// more than 24 keys: the table is rehashed on the way (32 -> 64 -> 128 -> 256 slots); rare, kept out of the hot code
__device__ __noinline__ void ua_fastutil_order_big(const uint8_t *in, int k, uint8_t *out, uint8_t *tab)
{
    int n = 32, size = 0;
    bool has_zero = false;
    for (int i = 0; i < n; i++) tab[i] = 0;
    for (int i = 0; i < k; i++) {
        const int key = in[i];
        if (key == 0) has_zero = true;
        else {
            int pos = (int)(ua_fu_mix(key) & (uint32_t)(n - 1));
            while (tab[pos] != 0) pos = (pos + 1) & (n - 1);
            tab[pos] = (uint8_t)(key + 1);
        }
        if (size++ >= n * 3 / 4) {                           // rehash(arraySize(size + 1, .75f)): the old table is walked downwards
            int nn = 2;
            const int need = (4 * (size + 1) + 2) / 3;       // ceil((size + 1) / .75)
            while (nn < need) nn <<= 1;
            // re-insert in place is not possible: move the keys out (downward walk order), clear, insert
            int cnt = 0;
            for (int j = n - 1; j >= 0; j--)
                if (tab[j] != 0) out[cnt++] = (uint8_t)(tab[j] - 1);
            for (int j = 0; j < nn; j++) tab[j] = 0;
            for (int j = 0; j < cnt; j++) {
                int pos = (int)(ua_fu_mix(out[j]) & (uint32_t)(nn - 1));
                while (tab[pos] != 0) pos = (pos + 1) & (nn - 1);
                tab[pos] = (uint8_t)(out[j] + 1);
            }
            n = nn;
        }
    }
    int o = 0;
    if (has_zero) out[o++] = 0;
    for (int j = n - 1; j >= 0; j--)
        if (tab[j] != 0) out[o++] = (uint8_t)(tab[j] - 1);
}

// single link (unreachable with the reference's thresholds: 3000 > 100): pairs by (cost, i, j), merged while cost <= cut; one lane
#ifndef UA_NI
#define UA_NI
#endif
template <int MAXN>
__device__ UA_NI void ua_single_link(UaShared<MAXN> &S, int m, int cut)
{
    for (int a = 0; a < m; a++) S.tmp[0][a] = (uint8_t)a;            // root of every leaf
    for (int s = 0; s <= cut; s++)
        for (int i = 0; i < m; i++)
            for (int j = i + 1; j < m; j++) {
                if (S.D[i][j] != s) continue;
                const int r1 = S.tmp[0][i], r2 = S.tmp[0][j];
                if (r1 == r2) continue;
                S.nxt[S.tail[r1]] = S.head[r2]; S.tail[r1] = S.tail[r2]; S.csize[r1] = (uint8_t)(S.csize[r1] + S.csize[r2]); S.csize[r2] = 0;
                for (int x = 0; x < m; x++) if (S.tmp[0][x] == r2) S.tmp[0][x] = (uint8_t)r1;
            }
}

// is the threshold graph a disjoint union of cliques?  (only then is the partition the same for every merge order)  Returns true when NOT.
template <int MAXN>
__device__ UA_NI bool ua_not_cluster_graph(const UaShared<MAXN> &S, int m, int cut, int lane)
{
    bool bad = false;
    for (int a = lane; a < m && !bad; a += 32) {
        const int ra = S.iwn[a];
        for (int b = 0; b < m && !bad; b++) {
            if (b == a || S.E[ra][S.iwn[b]] > cut) continue;
            const int rb = S.iwn[b];
            for (int c = 0; c < m; c++)
                if (c != a && c != b && ((S.E[ra][S.iwn[c]] <= cut) != (S.E[rb][S.iwn[c]] <= cut))) { bad = true; break; }
        }
    }
    return __any_sync(FULL, bad);
}

template <int MAXN>
__device__ void ua_job(UaShared<MAXN> &S, const int32_t *__restrict__ mat, int n, const slr_umi_assign_params P, int qv01,
                       slr_umi_assign_rec *__restrict__ rec, int lane)
{
    constexpr int NW = (MAXN + 31) / 32;                     // 32-lane rounds over the reads of a job
    // ---- the job's distances ------------------------------------------------------------------------------------------------------
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        S.E[i][j] = (uint8_t)(mat[idx] & 0xFF);              // getED(): (byte)(packed & 0xFFFFFF), 0 ... 5
    }
    __syncwarp();
    // ---- reads with a neighbour (always against umi_completelinkclusteringED) -----------------------------------------------------
    int m = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int i = w * 32 + lane;
        bool any = false;
        if (i < n)
            for (int j = 0; j < n; j++) any |= (j != i) && S.E[i][j] <= P.ed_complete;
        const unsigned b = __ballot_sync(FULL, any);
        if (any) S.iwn[m + __popc(b & ((1u << lane) - 1u))] = (uint8_t)i;
        m += __popc(b);
        if (i < n) S.clid[i] = 0xFF;
    }
    __syncwarp();
    if (m <= 1) return;                                      // ClusterOneHierarchical.java:L86
    const bool single = m > P.single_threshold;              // L79
    const int cut = single ? P.ed_single : P.ed_complete;
    // ---- leaves -------------------------------------------------------------------------------------------------------------------
    for (int a = lane; a < m; a += 32) {
        S.birth[a] = 0; S.head[a] = S.tail[a] = (uint8_t)a; S.nxt[a] = 0xFF; S.csize[a] = 1;
        const int ra = S.iwn[a];
        for (int b = 0; b < m; b++) S.D[a][b] = S.E[ra][S.iwn[b]];
    }
    __syncwarp();
    unsigned active[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) active[w] = (m - 32 * w >= 32) ? FULL : (m - 32 * w > 0 ? (1u << (m - 32 * w)) - 1u : 0u);
    bool tie_seen = false;
    if (!single) {
        // ---- complete link: poll the queue, merge, until the cheapest pair costs more than the cut -----------------------------------
        for (int step = 1; step < m; step++) {
            uint32_t best = 0xFFFFFFFFu;
            int best_a = 0;
            for (int a = 0; a < m; a++) {                    // pairs (a, b), a < b, both active; lane = b (mod 32)
                if (!((active[a >> 5] >> (a & 31)) & 1u)) continue;
                const int ba = S.birth[a];
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const int b = w * 32 + lane;
                    if (b > a && b < m && ((active[w] >> lane) & 1u)) {
                        const int bb = S.birth[b], gen = ba > bb ? ba : bb;
                        const uint32_t sub = gen == 0 ? (uint32_t)(a * m + b) : (uint32_t)S.R[a][b];
                        const uint32_t key = ((uint32_t)S.D[a][b] << 24) | ((uint32_t)(255 - gen) << 16) | (0xFFFFu - sub);
                        if (key < best) { best = key; best_a = a | (b << 8); }
                    }
                }
            }
            const uint32_t kmin = __reduce_min_sync(FULL, best);
            if (kmin == 0xFFFFFFFFu) break;
            const int score = (int)(kmin >> 24);
            if (score > cut) break;
            const int src = __ffs((int)__ballot_sync(FULL, best == kmin)) - 1;
            const int ab = __shfl_sync(FULL, best_a, src), a = ab & 0xFF, b = ab >> 8;
            const int ba = S.birth[a], bb = S.birth[b], gen = ba > bb ? ba : bb;
            const int d1 = (ba > bb) ? a : ((bb > ba) ? b : a), d2 = d1 == a ? b : a;   // mDendrogram1 = the younger cluster / the first leaf
            // a rival of the same generation at the same cost?  (pairs of one generation all contain the cluster born in that step)
            if (gen > 0) {
                bool rival = false;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const int c = w * 32 + lane;
                    if (c < m && ((active[w] >> lane) & 1u) && c != a && c != b && S.birth[c] < gen && S.D[d1][c] == score) rival = true;
                }
                tie_seen |= __any_sync(FULL, rival);
            }
            // new pairs (12, c) in the order of the ids of the pairs (2, c)
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const int c = w * 32 + lane;
                if (c < m) {
                    const int g2 = S.birth[d2] > S.birth[c] ? S.birth[d2] : S.birth[c];
                    const uint32_t sub = g2 == 0 ? (uint32_t)((d2 < c ? d2 : c) * m + (d2 < c ? c : d2)) : (uint32_t)S.R[d2][c];
                    const bool live = ((active[w] >> lane) & 1u) && c != d1 && c != d2;
                    S.idkey[c] = live ? (((uint32_t)g2 << 16) | sub) : 0xFFFFFFFFu;
                }
            }
            __syncwarp();
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const int c = w * 32 + lane;
                if (c < m && S.idkey[c] != 0xFFFFFFFFu) {
                    const uint32_t mine = S.idkey[c];
                    int rank = 0;
                    for (int c2 = 0; c2 < m; c2++) rank += S.idkey[c2] < mine;
                    const uint8_t nd = S.D[d1][c] > S.D[d2][c] ? S.D[d1][c] : S.D[d2][c];   // Math.max(dist1_3, dist2_3)
                    S.D[d1][c] = nd; S.D[c][d1] = nd;
                    S.R[d1][c] = (uint8_t)rank; S.R[c][d1] = (uint8_t)rank;
                }
            }
            if (lane == 0) {
                S.birth[d1] = (uint8_t)step;
                S.nxt[S.tail[d1]] = S.head[d2];              // members of dendrogram1, then of dendrogram2
                S.tail[d1] = S.tail[d2];
                S.csize[d1] = (uint8_t)(S.csize[d1] + S.csize[d2]);
            }
            active[d2 >> 5] &= ~(1u << (d2 & 31));
            __syncwarp();
        }
    } else if (lane == 0) {
        ua_single_link<MAXN>(S, m, cut);
    }
    __syncwarp();
    if (single) {
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const int a = w * 32 + lane;
            const unsigned alive = __ballot_sync(FULL, a < m && S.csize[a] > 0);
            active[w] = alive;
        }
    }
    // ---- clusters of more than one read, ordered like OneUmiCluster iterates them (sequential per cluster, lane 0) ------------------
    // cluster ordinal = position of its slot among the active slots
    int n_cl = 0, maxdepth = 0;
    bool chain_dep = false;
    if (lane == 0) {
        int o = 0;
        for (int s = 0; s < m; s++) {
            if (!((active[s >> 5] >> (s & 31)) & 1u)) continue;
            const int k = S.csize[s];
            if (k <= 1) continue;
            int x = S.head[s];
            if (n <= 16) {
                // every index is below the smallest HashSet table (16 buckets): both sets iterate in ascending order, no chain
                uint32_t mem = 0;
                for (int i = 0; i < k; i++) { mem |= 1u << S.iwn[x]; x = S.nxt[x]; }
                for (int i = 0; i < k; i++) { const int b = __ffs((int)mem) - 1; mem &= mem - 1u; S.tmp[1][i] = (uint8_t)b; }
            } else {
                for (int i = 0; i < k; i++) { S.tmp[0][i] = (uint8_t)x; x = S.nxt[x]; }                 // memberSet(): reduced indices
                chain_dep |= ua_jdk_order(S.tmp[0], k, S.tmp[1]);
                for (int i = 0; i < k; i++) S.tmp[0][i] = S.iwn[S.tmp[1][i]];                          // transformIndices_AndRemoveSingletons
                chain_dep |= ua_jdk_order(S.tmp[0], k, S.tmp[1]);
            }
            ua_fastutil_order(S.tmp[1], k, &S.it[o], S.tab);                                           // toCollection(OneUmiCluster::new)
            S.cl_start[n_cl] = (uint8_t)o; S.cl_len[n_cl] = (uint8_t)k;
            for (int i = 0; i < k; i++) S.clid[S.it[o + i]] = (uint8_t)n_cl;
            o += k;
            if (k > maxdepth) maxdepth = k;
            n_cl++;
        }
    }
    n_cl = __shfl_sync(FULL, n_cl, 0);
    maxdepth = __shfl_sync(FULL, maxdepth, 0);
    chain_dep = __shfl_sync(FULL, (int)chain_dep, 0) != 0;
    __syncwarp();
    if (n_cl == 0) return;
    bool unpinned = false;
    if (tie_seen) unpinned = ua_not_cluster_graph<MAXN>(S, m, cut, lane) || chain_dep;
    // ---- depth rule, centres, mean shift: lane = cluster --------------------------------------------------------------------------
    int n_list = 0;
    for (int c0 = 0; c0 < n_cl; c0 += 32) {
        const int c = c0 + lane;
        bool pass = false;
        if (c < n_cl) {
            const int k = S.cl_len[c];
            const uint8_t *it = &S.it[S.cl_start[c]];
            pass = k * P.fold_depth > maxdepth;              // ClusterOneHierarchical.java:L123
            int center = 0xFF, off_mean = 0;
            if (pass) {
                if (k == 2) center = qv01 ? it[0] : it[1];   // OneUmiCluster.java:L52-L54
                else {
                    int bests = 0x7FFFFFFF;
                    for (int i = 0; i < k; i++) {
                        int s = 0;
                        for (int j = 0; j < k; j++) { const int e = S.E[it[i]][it[j]]; s += (j != i) ? e * e : 0; }
                        if (s < bests) { bests = s; center = it[i]; }
                    }
                }
                int sum = 0;
                for (int i = 0; i < k; i++)
                    if (it[i] != center) sum += ua_pos1_offset(mat[center * n + it[i]]);
                const int cnt = k - 1;                       // Math.round(sum / cnt) = floor((2 sum + cnt) / (2 cnt))
                const int num = 2 * sum + cnt, den = 2 * cnt;
                off_mean = num >= 0 ? num / den : -((-num + den - 1) / den);
            }
            S.cl_center[c] = (uint8_t)center; S.cl_off[c] = (int8_t)off_mean;
        }
        n_list += __popc(__ballot_sync(FULL, pass));
    }
    __syncwarp();
    // ---- per read ------------------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int x = w * 32 + lane;
        if (x >= n) continue;
        slr_umi_assign_rec r;
        r.center = -1; r.u1 = 0; r.u2 = -1; r.pos2 = 0; r.offset_center_mean = 0; r.flags = unpinned ? SLR_UA_TIE_UNPIN : 0; r.cluster_size = 0;
        r.n_clusters = n_list;
        const int c = S.clid[x];
        if (c != 0xFF) {
            const int center = S.cl_center[c];
            r.cluster_size = S.cl_len[c];
            if (center == 0xFF) r.flags |= SLR_UA_SKIPPED;   // flagDontUMIassignRecords (ClusterOneBase.java:L57, L71)
            else {
                const int32_t cell = mat[center * n + x];
                r.center = center; r.flags |= SLR_UA_ASSIGNED; r.offset_center_mean = S.cl_off[c];
                r.u1 = (int8_t)(cell & 0xFF);                // distanceNonReducedSet(center, index) (ClusterOneBase.java:L156)
                r.pos2 = (int8_t)ua_pos2_code(cell);         // L133
                if (n_list > 1) {                            // L161-L164
                    int best = 0x7F;
                    for (int y = 0; y < n; y++)
                        if (S.clid[y] != c && S.E[x][y] < best) best = S.E[x][y];
                    r.u2 = (int8_t)(best == 0x7F ? -1 : best);
                }
            }
        }
        rec[x] = r;
    }
}

// ---- kernels ---------------------------------------------------------------------------------------------------------------------------
// records of every read: default values; reads of jobs above max_hier are flagged SLR_UA_DEEP.  The first read of a job files the job in
// the list of its class: lists[0 .. n_jobs) small jobs (2 ... 32 reads), lists[n_jobs .. 2 n_jobs) large jobs (33 ... max_hier),
// lists[2 n_jobs .. 3 n_jobs) / [3 n_jobs .. 4 n_jobs) deep jobs of up to / above SLR_UA_DEEP_SMALL reads (umi_assign_deep.cu) together with the
// offset of their working arrays in the deep arena — a deep job that does not fit the arena stays flagged SLR_UA_DEEP only.
__global__ void __launch_bounds__(256) umi_assign_init(const long long *__restrict__ joff, long long n_jobs, long long n_reads,
                                                        const int32_t *__restrict__ rowjob, int max_hier, slr_umi_assign_rec *__restrict__ rec,
                                                        int32_t *__restrict__ lists, unsigned int *__restrict__ counts,
                                                        long long *__restrict__ deep_off, long long deep_cap_words,
                                                        const int32_t *__restrict__ mat, const long long *__restrict__ ooff,
                                                        const slr_umi_assign_params P, const uint8_t *__restrict__ job_qv01)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        long long j = -1;
        if (rowjob) j = rowjob[r];
        else if (r >= joff[0] && r < joff[n_jobs]) {
            long long lo = 0, hi = n_jobs;
            while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (joff[mid] <= r) lo = mid; else hi = mid; }
            j = lo;
        }
        slr_umi_assign_rec d;
        d.center = -1; d.u1 = 0; d.u2 = -1; d.pos2 = 0; d.offset_center_mean = 0; d.flags = 0; d.cluster_size = 0; d.n_clusters = 0;
        if (j >= 0) {
            const long long r0 = joff[j], n = joff[j + 1] - r0;
            if (n > max_hier) {
                d.flags = SLR_UA_DEEP;
                if (r == r0 && deep_cap_words > 0) {
                    const long long need = slr_umi_assign_deep_words(n);
                    const long long off = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(counts + 4), (unsigned long long)need);
                    if (off + need <= deep_cap_words) {
                        const int cls = n <= SLR_UA_DEEP_SMALL ? 2 : 3;
                        const unsigned int k = atomicAdd(&counts[cls], 1u);
                        lists[(long long)cls * n_jobs + k] = (int32_t)j;
                        deep_off[(long long)(cls - 2) * n_jobs + k] = off;
                    }
                }
            } else if (n == 2) {
                // a job of two reads (a quarter of all jobs that have anything to cluster) needs no warp: one pair, no ties, the HashSet and
                // fastutil orders of {0, 1} are (0, 1); each of its two reads writes its own record
                const int32_t *M = mat + ooff[j];
                const int32_t m01 = M[1], m10 = M[2];
                const int edc = P.ed_complete, e01 = (int)(int8_t)(m01 & 0xFF), e10 = (int)(int8_t)(m10 & 0xFF);
                const int nn = (e01 <= edc) + (e10 <= edc);                       // reads with a neighbour (DistanceMatrix.java:L87-L90)
                const int cut = nn > P.single_threshold ? P.ed_single : edc;
                if (nn == 2 && e01 <= cut) {
                    if (2 * P.fold_depth > 2) {
                        const int center = (job_qv01 && job_qv01[j]) ? 0 : 1, x = (int)(r - r0);
                        const int32_t cell = M[center * 2 + x];
                        d.center = center; d.flags = SLR_UA_ASSIGNED; d.cluster_size = 2; d.n_clusters = 1;
                        d.offset_center_mean = (int8_t)ua_pos1_offset(M[center * 2 + (1 - center)]);
                        d.u1 = (int8_t)(int)(int8_t)(cell & 0xFF); d.pos2 = (int8_t)ua_pos2_code(cell);
                    } else { d.flags = SLR_UA_SKIPPED; d.cluster_size = 2; }
                }
            } else if (n == 3 && 3 <= P.single_threshold) {
                // three reads, complete link: three pairs in the queue (least cost first, among equals the pair offered last: (1,2), (0,2), (0,1)),
                // at most two merges, and the only pair a merge creates has no rival, so no tie can matter.  HashSet order of any subset of
                // {0, 1, 2} is ascending; fastutil iterates key 0 first, then slot 28 (key 2) before slot 14 (key 1).
                const int32_t *M = mat + ooff[j];
                const int edc = P.ed_complete, x = (int)(r - r0);
                int e[3][3];
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++) e[a][b] = (int)(int8_t)(M[a * 3 + b] & 0xFF);
                const bool nb0 = e[0][1] <= edc || e[0][2] <= edc, nb1 = e[1][0] <= edc || e[1][2] <= edc, nb2 = e[2][0] <= edc || e[2][1] <= edc;
                const int m = (int)nb0 + (int)nb1 + (int)nb2;
                unsigned members = 0;                                            // bit i = read i is in THE cluster (there is at most one)
                if (m == 2) {
                    const int a = nb0 ? 0 : 1, b = nb2 ? 2 : 1;
                    if (e[a][b] <= edc) members = (1u << a) | (1u << b);
                } else if (m == 3) {
                    const int d01 = e[0][1], d02 = e[0][2], d12 = e[1][2];
                    int fa, fb, fz, fd;                                          // first poll: least cost, the largest pair id among equals
                    if (d12 <= d02 && d12 <= d01) { fa = 1; fb = 2; fz = 0; fd = d12; }
                    else if (d02 <= d01) { fa = 0; fb = 2; fz = 1; fd = d02; }
                    else { fa = 0; fb = 1; fz = 2; fd = d01; }
                    if (fd <= edc) {
                        members = (1u << fa) | (1u << fb);
                        const int da = fa < fz ? e[fa][fz] : e[fz][fa], db = fb < fz ? e[fb][fz] : e[fz][fb];     // the queue holds upper-triangle costs
                        if ((da > db ? da : db) <= edc) members = 7u;
                    }
                }
                if (members) {
                    const int k = __popc(members);
                    if (k * P.fold_depth > k) {
                        int it[3], c = 0;                                        // OneUmiCluster iteration order
                        if (members & 1u) it[c++] = 0;
                        if (members & 4u) it[c++] = 2;
                        if (members & 2u) it[c++] = 1;
                        int center;
                        if (k == 2) center = (job_qv01 && job_qv01[j]) ? it[0] : it[1];
                        else {
                            int best = 0x7FFFFFFF; center = it[0];
#pragma unroll
                            for (int q = 0; q < 3; q++) {
                                const int a = it[q];
                                int sum = 0;
#pragma unroll
                                for (int b = 0; b < 3; b++) if (b != a) sum += e[a][b] * e[a][b];
                                if (sum < best) { best = sum; center = a; }
                            }
                        }
                        d.n_clusters = 1;
                        if ((members >> x) & 1u) {
                            int sum = 0;
                            for (int q = 0; q < k; q++) if (it[q] != center) sum += ua_pos1_offset(M[center * 3 + it[q]]);
                            const int32_t cell = M[center * 3 + x];
                            d.center = center; d.flags = SLR_UA_ASSIGNED; d.cluster_size = (uint16_t)k;
                            d.offset_center_mean = (int8_t)(k == 2 ? sum : ((sum + 1) >> 1));      // Math.round of the mean of one or two shifts
                            d.u1 = (int8_t)(int)(int8_t)(cell & 0xFF); d.pos2 = (int8_t)ua_pos2_code(cell);
                        }
                    } else if ((members >> x) & 1u) { d.flags = SLR_UA_SKIPPED; d.cluster_size = (uint16_t)k; }
                }
            } else if (r == r0 && n >= 2) {
                const int cls = n <= 32 ? 0 : 1;
                lists[(long long)cls * n_jobs + atomicAdd(&counts[cls], 1u)] = (int32_t)j;
            }
        }
        rec[r] = d;
    }
}

template <int MAXN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) umi_assign_kernel(const int32_t *__restrict__ mat, const long long *__restrict__ joff,
                                                                const long long *__restrict__ ooff, const slr_umi_assign_params P,
                                                                const uint8_t *__restrict__ job_qv01, slr_umi_assign_rec *__restrict__ rec,
                                                                const int32_t *__restrict__ list, const unsigned int *__restrict__ count)
{
    __shared__ UaShared<MAXN> sm[WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned int total = *count;
    for (unsigned int k = blockIdx.x * WARPS + wib; k < total; k += gridDim.x * WARPS) {
        const long long j = list[k];
        const long long r0 = joff[j];
        const int n = (int)(joff[j + 1] - r0);
        ua_job<MAXN>(sm[wib], mat + ooff[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0, lane);
        __syncwarp();
    }
}

}  // namespace

// scratch: counts (8 words: 4 list lengths, the 64-bit arena cursor) | 4 job lists | 2 lists of arena offsets | the deep arena
size_t slr_umi_assign_scratch(long long n_jobs, long long deep_words)
{
    return 32 + (size_t)(4 * n_jobs + 4) * 4 + (size_t)(2 * n_jobs + 2) * 8 + (size_t)(deep_words > 0 ? deep_words : 0) * 4 + 16;
}

cudaError_t slr_launch_umi_assign(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                  long long n_reads, const slr_umi_assign_params &P, const uint8_t *d_job_qv01, const int32_t *d_rowjob,
                                  slr_umi_assign_rec *d_rec, void *d_scratch, size_t scratch_bytes, cudaStream_t stream)
{
    if (n_reads <= 0 || n_jobs <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (scratch_bytes < slr_umi_assign_scratch(n_jobs, 0)) return cudaErrorInvalidValue;
    unsigned int *counts = reinterpret_cast<unsigned int *>(d_scratch);
    int32_t *lists = reinterpret_cast<int32_t *>(counts + 8);
    long long *deep_off = reinterpret_cast<long long *>(lists + 4 * n_jobs + 4);
    int *arena = reinterpret_cast<int *>(deep_off + 2 * n_jobs + 2);
    const long long deep_cap = P.deep ? (long long)((scratch_bytes - slr_umi_assign_scratch(n_jobs, 0)) / 4) & ~1ll : 0;
    cudaError_t e = cudaMemsetAsync(counts, 0, 32, stream);
    if (e != cudaSuccess) return e;
    long long g0 = (n_reads + 255) / 256;
    if (g0 > (long long)sms * 16) g0 = (long long)sms * 16;
    slr_umi_assign_params Q = P;
    if (Q.max_hier > 100) Q.max_hier = 100;
    umi_assign_init<<<(unsigned)g0, 256, 0, stream>>>(d_job_offsets, n_jobs, n_reads, d_rowjob, Q.max_hier, d_rec, lists, counts, deep_off, deep_cap, d_mat,
                                                        d_out_offsets, Q, d_job_qv01);
    // persistent grids, one wave; warps without a job return at once
    long long gs = (n_jobs + 7) / 8;
    if (gs > (long long)sms * 6) gs = (long long)sms * 6;
    umi_assign_kernel<32, 8><<<(unsigned)gs, 256, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, Q, d_job_qv01, d_rec, lists, counts);
    long long gl = n_jobs < (long long)sms * 4 ? n_jobs : (long long)sms * 4;
    umi_assign_kernel<100, 1><<<(unsigned)gl, 32, 0, stream>>>(d_mat, d_job_offsets, d_out_offsets, Q, d_job_qv01, d_rec, lists + n_jobs,
                                                             counts + 1);
    e = cudaGetLastError();
    if (e != cudaSuccess || deep_cap <= 0) return e;
    return slr_launch_umi_assign_deep(d_mat, d_job_offsets, d_out_offsets, Q, d_job_qv01, d_rec, lists + 2 * n_jobs, deep_off, counts + 2,
                                      lists + 3 * n_jobs, deep_off + n_jobs, counts + 3, arena, n_jobs, stream);
}
