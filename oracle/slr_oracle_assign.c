/*
 * slr_oracle_assign.c — CPU ORACLE (test infrastructure, NOT product code): clustering of one (cell, region) job on its packed
 * distance matrix and the per-read "UMI assignment" derived from the clusters.
 *
 * Restated from the bytecode (F! = Jar/NanoporeBC_UMI_finder-2.1.jar, A! = Jar/lib/Aliasi_ClusteringLib-1.0.jar):
 *   UmiClustering$Submitter dispatch            F!…/clustering/UmiClustering$Submitter.class (UmiClustering.java:L239-L261): n <= 100 -> ClusterOneHierarchical,
 *                                               n > 100 -> ClusterOne_MyClustering; CLUSTERHOW is the constant DECIDEONCOMPLEXITY (…java:L50), so
 *                                               the pre-grouping branch (L242-L244: size > 1000 AND ALIASI) is unreachable: wasPregrouped is false.
 *   ClusterOneHierarchical.call                 F!…/clustering/ClusterOneHierarchical.class (ClusterOneHierarchical.java:L61-L217)
 *   DistanceMatrix (indicesWithNeighbors, distance, transformIndices_AndRemoveSingletons)   F!com/rw/clustering/DistanceMatrix.class (…java:L77-L90, L145, L158, L169)
 *   CompleteLinkClusterer.hierarchicalCluster   A!com/aliasi/cluster/CompleteLinkClusterer.class (CompleteLinkClusterer.java:L146-L237)
 *   SingleLinkClusterer.hierarchicalCluster     A!com/aliasi/cluster/SingleLinkClusterer.class (SingleLinkClusterer.java:L198-L268)
 *   BoundedPriorityQueue / EntryComparator      A!com/aliasi/util/BoundedPriorityQueue.class (BoundedPriorityQueue.java:L131-L153, L342-L346, L372-L379, L458-L464)
 *   Dendrogram.partitionDistance                A!com/aliasi/cluster/Dendrogram.class (Dendrogram.java:L205-L215), LinkDendrogram (LinkDendrogram.java:L85-L94, L129-L130)
 *   OneUmiCluster.setClusterCenterNotPreGrouped F!com/rw/clustering/OneUmiCluster.class (OneUmiCluster.java:L49-L65)
 *   ClusterOneBase.setSamflagsAndStatsForClustered   F!…/clustering/ClusterOneBase.class (ClusterOneBase.java:L118-L168)
 *
 * Containers whose iteration order reaches the result are modelled explicitly:
 *   java.util.HashSet<Integer>  (JDK HashMap: table of 16 doubling above a load of .75, bucket = hash & (cap - 1), hash(Integer) = value,
 *                                chains in insertion order, resize keeps the relative order)
 *   fastutil IntOpenHashSet     (OneUmiCluster's superclass; 8.2.2 layout as published: 32 slots for <= 24 keys, slot = mix(k) & mask,
 *                                linear probing, key 0 outside the table and iterated first, then the slots from the last to the first)
 *   ObjectToSet's HashSet<PairScore> is keyed by IDENTITY hash codes: the reference's own iteration order is JVM-run dependent.  The
 *   canonical order used here (and by the GPU kernel) is the creation order of the pairs; a job whose result can depend on that order is
 *   flagged ORC_UA_TIE_UNPIN.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "slr_oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

static int ed_of(int32_t packed) { return (int)(int8_t)(packed & 0xFFFFFF); }            /* BestEditDistance.getED: iand, i2b (L382) */
static int pos1_offset(int32_t p) { return (p & 0x08000000) ? -1 : (p & 0x10000000) ? 0 : (p & 0x20000000) ? 1 : 0; }   /* getPos1().getOffSet() */
static int pos2_code(int32_t p) { return (p & 0x01000000) ? 0 : (p & 0x02000000) ? 1 : (p & 0x04000000) ? 2 : 1; }      /* getPos2(): MINUSONE, ZERO, PLUSONE */

/* ---- java.util.HashSet<Integer> iteration order --------------------------------------------------------------------------- */
static int jdk_cap_for(int size)
{
    int cap = 16;
    while (size > cap * 3 / 4) cap <<= 1;          /* resize when ++size > threshold (= .75 * cap) */
    return cap;
}
/* in: elements in insertion order; out: iteration order.  Returns 1 when two elements share a bucket (chain order = insertion order). */
static int jdk_hashset_order(const int *in, int k, int *out)
{
    const int cap = jdk_cap_for(k);
    int o = 0, collide = 0;
    for (int b = 0; b < cap; b++) {
        int cnt = 0;
        for (int i = 0; i < k; i++)
            if ((in[i] & (cap - 1)) == b) { out[o++] = in[i]; cnt++; }
        if (cnt > 1) collide = 1;
    }
    return collide;
}

/* ---- fastutil IntOpenHashSet iteration order -------------------------------------------------------------------------------- */
static uint32_t fu_mix(int k) { const uint32_t h = (uint32_t)k * 0x9E3779B9u; return h ^ (h >> 16); }     /* HashCommon.mix */
static int fu_array_size(int expected)      /* HashCommon.arraySize(expected, .75f) = max(2, nextPowerOfTwo(ceil(expected / f))) */
{
    long need = (long)ceil(expected / 0.75f), n = 2;
    while (n < need) n <<= 1;
    return (int)n;
}
static void fastutil_intset_order(const int *in, int k, int *out)
{
    int n = 32, size = 0, has_zero = 0;
    int *tab = (int *)calloc((size_t)n, sizeof(int));
    for (int i = 0; i < k; i++) {
        const int key = in[i];
        if (key == 0) { if (has_zero) continue; has_zero = 1; }
        else {
            int pos = (int)(fu_mix(key) & (uint32_t)(n - 1)), dup = 0;
            while (tab[pos] != 0) { if (tab[pos] == key) { dup = 1; break; } pos = (pos + 1) & (n - 1); }
            if (dup) continue;
            tab[pos] = key;
        }
        if (size++ >= n * 3 / 4) {                                                      /* maxFill(n, .75f) */
            const int nn = fu_array_size(size + 1);
            int *nt = (int *)calloc((size_t)nn, sizeof(int));
            for (int j = n - 1; j >= 0; j--)                                            /* rehash walks the old table downwards */
                if (tab[j] != 0) {
                    int pos = (int)(fu_mix(tab[j]) & (uint32_t)(nn - 1));
                    while (nt[pos] != 0) pos = (pos + 1) & (nn - 1);
                    nt[pos] = tab[j];
                }
            free(tab); tab = nt; n = nn;
        }
    }
    int o = 0;
    if (has_zero) out[o++] = 0;
    for (int j = n - 1; j >= 0; j--) if (tab[j] != 0) out[o++] = tab[j];
    free(tab);
}

/* ---- dendrograms ------------------------------------------------------------------------------------------------------------ */
typedef struct { int left, right, leaf, parent; double score; } dnode;     /* leaf >= 0: LeafDendrogram(object), score 0 */
typedef struct { int d1, d2; double score; long id; int in_queue; } pscore; /* PairScore + its BoundedPriorityQueue entry id */

static int deref(const dnode *D, int x) { while (D[x].parent >= 0) x = D[x].parent; return x; }     /* Dendrogram.dereference */
static int members(const dnode *D, int x, int *out, int o)                                           /* addMembers: left, then right */
{
    if (D[x].leaf >= 0) { out[o++] = D[x].leaf; return o; }
    o = members(D, D[x].left, out, o);
    return members(D, D[x].right, out, o);
}

/* CompleteLinkClusterer.hierarchicalCluster over elements 0..m-1 (the HashSet {0..m-1} iterates ascending: every value is below the
 * table size).  Stops once the cheapest pair costs more than max_ed: every later link is above the cut of partitionDistance(max_ed).
 * Returns the number of dendrogram nodes; *tie_seen = some poll at cost <= max_ed had a same-cost rival created in the same merge step
 * (their relative order in the queue follows the identity-hash order of ObjectToSet's HashSet). */
static int complete_link(const int *dist, int m, int max_ed, dnode *D, int *tie_seen)
{
    int nd = m, np = 0, cap = m * (m - 1) / 2 + m * m + 8;
    long next_id = 0;
    pscore *P = (pscore *)malloc((size_t)cap * sizeof(pscore));
    /* index: per dendrogram node the pairs it is part of, in insertion order */
    int **idx = (int **)calloc((size_t)(2 * m), sizeof(int *)), *idx_n = (int *)calloc((size_t)(2 * m), sizeof(int));
    int *born = (int *)calloc((size_t)cap, sizeof(int));                                /* merge step that created the pair (0 = initial) */
    for (int i = 0; i < 2 * m; i++) idx[i] = (int *)malloc((size_t)(2 * m + 4) * sizeof(int));
    for (int i = 0; i < m; i++) { D[i].left = D[i].right = -1; D[i].leaf = i; D[i].parent = -1; D[i].score = 0.0; }
    for (int i = 0; i < m; i++)
        for (int j = i + 1; j < m; j++) {                                               /* L169-L180 */
            P[np].d1 = i; P[np].d2 = j; P[np].score = (double)dist[i * m + j]; P[np].id = next_id++; P[np].in_queue = 1; born[np] = 0;
            idx[i][idx_n[i]++] = np; idx[j][idx_n[j]++] = np;
            np++;
        }
    *tie_seen = 0;
    int step = 0;
    for (;;) {
        int best = -1;                                                                  /* poll(): least cost, among equals the LARGEST id (L458-L464) */
        for (int p = 0; p < np; p++)
            if (P[p].in_queue && (best < 0 || P[p].score < P[best].score || (P[p].score == P[best].score && P[p].id > P[best].id))) best = p;
        if (best < 0) break;
        if (P[best].score > (double)max_ed) break;
        if (born[best] > 0)
            for (int p = 0; p < np; p++)
                if (p != best && P[p].in_queue && P[p].score == P[best].score && born[p] == born[best]) *tie_seen = 1;
        step++;
        P[best].in_queue = 0;
        const int d1 = deref(D, P[best].d1), d2 = deref(D, P[best].d2), d12 = nd++;     /* L186-L189 */
        D[d12].left = d1; D[d12].right = d2; D[d12].leaf = -1; D[d12].parent = -1; D[d12].score = P[best].score;
        D[d1].parent = d12; D[d2].parent = d12;
        idx[d12] = idx[d12] ? idx[d12] : (int *)malloc((size_t)(2 * m + 4) * sizeof(int));
        double *buf = (double *)malloc((size_t)(2 * m) * sizeof(double));               /* distanceBuf (L193) */
        char *has = (char *)calloc((size_t)(2 * m), 1);
        for (int q = 0; q < idx_n[d1]; q++) {                                           /* L195-L205 */
            const int p = idx[d1][q];
            if (p < 0) continue;
            P[p].in_queue = 0;
            const int d3 = (P[p].d1 == d1) ? P[p].d2 : P[p].d1;
            for (int r = 0; r < idx_n[d3]; r++) if (idx[d3][r] == p) idx[d3][r] = -1;
            buf[d3] = P[p].score; has[d3] = 1;
        }
        idx_n[d1] = 0;
        for (int q = 0; q < idx_n[d2]; q++) {                                           /* L208-L225: canonical order = insertion order */
            const int p = idx[d2][q];
            if (p < 0) continue;
            P[p].in_queue = 0;
            const int d3 = (P[p].d1 == d2) ? P[p].d2 : P[p].d1;
            for (int r = 0; r < idx_n[d3]; r++) if (idx[d3][r] == p) idx[d3][r] = -1;
            if (!has[d3]) continue;
            P[np].d1 = d12; P[np].d2 = d3; P[np].score = buf[d3] > P[p].score ? buf[d3] : P[p].score;     /* Math.max (L220) */
            P[np].id = next_id++; P[np].in_queue = 1; born[np] = step;
            idx[d12][idx_n[d12]++] = np; idx[d3][idx_n[d3]++] = np;
            np++;
        }
        idx_n[d2] = 0;
        free(buf); free(has);
    }
    for (int i = 0; i < 2 * m; i++) free(idx[i]);
    free(idx); free(idx_n); free(P); free(born);
    return nd;
}

typedef struct { double score; int i, j, seq; } slpair;
static int slpair_cmp(const void *a, const void *b)       /* ScoredObject.comparator() under the stable Arrays.sort (TimSort) */
{
    const slpair *x = (const slpair *)a, *y = (const slpair *)b;
    if (x->score != y->score) return x->score < y->score ? -1 : 1;
    return x->seq - y->seq;
}
/* SingleLinkClusterer.hierarchicalCluster (L205-L268): pairs sorted by cost, merged while cost <= maxDistance */
static int single_link(const int *dist, int m, int max_ed, dnode *D)
{
    int nd = m, np = 0;
    slpair *P = (slpair *)malloc((size_t)(m * (m - 1) / 2 + 1) * sizeof(slpair));
    for (int i = 0; i < m; i++) { D[i].left = D[i].right = -1; D[i].leaf = i; D[i].parent = -1; D[i].score = 0.0; }
    for (int i = 0; i < m; i++)
        for (int j = i + 1; j < m; j++) { P[np].score = (double)dist[i * m + j]; P[np].i = i; P[np].j = j; P[np].seq = np; np++; }
    qsort(P, (size_t)np, sizeof(slpair), slpair_cmp);
    int clusters = m;
    for (int p = 0; p < np && clusters > 1; p++) {
        if (P[p].score > (double)max_ed) break;
        const int d1 = deref(D, P[p].i), d2 = deref(D, P[p].j);
        if (d1 == d2) continue;
        D[nd].left = d1; D[nd].right = d2; D[nd].leaf = -1; D[nd].parent = -1; D[nd].score = P[p].score;
        D[d1].parent = nd; D[d2].parent = nd;
        nd++; clusters--;
    }
    free(P);
    return nd;
}

/* ClusterOneHierarchical.call for one job (n reads, n x n packed matrix). */
void orc_umi_assign_hier(const int32_t *matrix, int64_t n64, const orc_assign_params *P, int qv01, orc_assign_rec *rec)
{
    const int n = (int)n64;
    for (int i = 0; i < n; i++) { memset(&rec[i], 0, sizeof(rec[i])); rec[i].center = -1; rec[i].u2 = -1; }
    if (n < 2) return;
    /* DistanceMatrix.generateIndicesWithNeighbours (L87-L90): always against umi_completelinkclusteringED */
    int *iwn = (int *)malloc((size_t)n * sizeof(int)), m = 0;
    for (int i = 0; i < n; i++) {
        int any = 0;
        for (int j = 0; j < n; j++) if (i != j && ed_of(matrix[(size_t)i * n + j]) <= P->ed_complete) any = 1;
        if (any) iwn[m++] = i;
    }
    if (m <= 1) { free(iwn); return; }                                                   /* L86 */
    const int single = m > P->single_threshold;                                          /* L79 */
    const int cut = single ? P->ed_single : P->ed_complete;                              /* L83-L84, L101 */
    int *dist = (int *)malloc((size_t)m * m * sizeof(int));
    for (int a = 0; a < m; a++)
        for (int b = 0; b < m; b++) dist[a * m + b] = ed_of(matrix[(size_t)iwn[a] * n + iwn[b]]);      /* DistanceMatrix.distance (L158) */
    dnode *D = (dnode *)malloc((size_t)(2 * m) * sizeof(dnode));
    int tie_seen = 0;
    const int nd = single ? single_link(dist, m, cut, D) : complete_link(dist, m, cut, D, &tie_seen);
    /* is the threshold graph a disjoint union of cliques?  (then every merge order ends in the same partition) */
    int cluster_graph = 1;
    for (int a = 0; a < m && cluster_graph; a++)
        for (int b = 0; b < m && cluster_graph; b++)
            if (a != b && dist[a * m + b] <= cut)
                for (int c = 0; c < m; c++)
                    if (c != a && c != b && (dist[a * m + c] <= cut) != (dist[b * m + c] <= cut)) { cluster_graph = 0; break; }
    /* partitionDistance(cut): the roots left over are the maximal subtrees with cost <= cut (later links cost more); size > 1 (L101);
     * transformIndices_AndRemoveSingletons (L104, DistanceMatrix.java:L145) */
    int *cl_of = (int *)malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; i++) cl_of[i] = -1;
    int n_cl = 0, maxdepth = 0, chain_dep = 0;
    int *cl_size = (int *)calloc((size_t)m, sizeof(int)), **cl_it = (int **)calloc((size_t)m, sizeof(int *));
    int *tmpA = (int *)malloc((size_t)m * sizeof(int)), *tmpB = (int *)malloc((size_t)m * sizeof(int));
    for (int x = 0; x < nd; x++) {
        if (D[x].parent >= 0) continue;
        const int k = members(D, x, tmpA, 0);
        if (k <= 1) continue;
        chain_dep |= jdk_hashset_order(tmpA, k, tmpB);                                   /* memberSet(): HashSet of reduced indices */
        for (int i = 0; i < k; i++) tmpA[i] = iwn[tmpB[i]];                              /* stream().map(k -> indicesWithNeighbors.get(k)) */
        chain_dep |= jdk_hashset_order(tmpA, k, tmpB);                                   /* collect(toSet()) */
        cl_it[n_cl] = (int *)malloc((size_t)k * sizeof(int));
        fastutil_intset_order(tmpB, k, cl_it[n_cl]);                                     /* toCollection(OneUmiCluster::new) (L129) */
        cl_size[n_cl] = k;
        if (k > maxdepth) maxdepth = k;                                                  /* L118 */
        n_cl++;
    }
    const int unpinned = tie_seen && (!cluster_graph || chain_dep);
    /* depth rule (L121-L127) */
    int n_list = 0;
    for (int c = 0; c < n_cl; c++) if (cl_size[c] * P->fold_depth > maxdepth) n_list++;
    for (int c = 0; c < n_cl; c++) {
        const int k = cl_size[c], *it = cl_it[c];
        if (!(k * P->fold_depth > maxdepth)) {                                           /* flagDontUMIassignRecords (ClusterOneBase.java:L57, L71) */
            for (int i = 0; i < k; i++) { rec[it[i]].flags |= ORC_UA_SKIPPED; rec[it[i]].cluster_size = (uint16_t)k; }
            continue;
        }
        int center;
        if (k == 2) center = qv01 ? it[0] : it[1];                                       /* OneUmiCluster.java:L52-L54 */
        else {                                                                           /* L57-L63: least sum of squared distances, first wins */
            long bests = -1; center = it[0];
            for (int i = 0; i < k; i++) {
                long s = 0;
                for (int j = 0; j < k; j++) if (it[j] != it[i]) { const int e = ed_of(matrix[(size_t)it[i] * n + it[j]]); s += (long)pow((double)e, 2.0); }
                if (bests < 0 || s < bests) { bests = s; center = it[i]; }
            }
        }
        long sum = 0, cnt = 0;                                                           /* ClusterOneHierarchical.java:L143-L147 */
        for (int i = 0; i < k; i++) if (it[i] != center) { sum += pos1_offset(matrix[(size_t)center * n + it[i]]); cnt++; }
        const int off_mean = (int)floor((double)sum / (double)cnt + 0.5);                /* Math.round(average) */
        for (int i = 0; i < k; i++) {                                                    /* L178-L195 -> setSamflagsAndStatsForClustered */
            const int x = it[i];
            orc_assign_rec *r = &rec[x];
            r->center = center; r->flags |= ORC_UA_ASSIGNED; r->cluster_size = (uint16_t)k; r->off_mean = (int8_t)off_mean;
            r->u1 = (int8_t)ed_of(matrix[(size_t)center * n + x]);                        /* distanceNonReducedSet(center, index) (L156) */
            r->pos2 = (int8_t)pos2_code(matrix[(size_t)center * n + x]);                  /* L133 */
            cl_of[x] = c;
        }
    }
    for (int c = 0; c < n_cl; c++) {                                                     /* U2 (L161-L164): cluster_list.size() > 1, min over y outside thisCluster */
        if (!(cl_size[c] * P->fold_depth > maxdepth)) continue;
        for (int i = 0; i < cl_size[c]; i++) {
            const int x = cl_it[c][i];
            if (n_list > 1) {
                int best = -1;
                for (int y = 0; y < n; y++) {
                    int inside = 0;
                    for (int j = 0; j < cl_size[c]; j++) if (cl_it[c][j] == y) inside = 1;
                    if (inside) continue;
                    const int e = ed_of(matrix[(size_t)x * n + y]);
                    if (best < 0 || e < best) best = e;
                }
                rec[x].u2 = (int8_t)best;
            }
        }
    }
    for (int i = 0; i < n; i++) {
        rec[i].n_clusters = n_list;
        if (unpinned) rec[i].flags |= ORC_UA_TIE_UNPIN;
    }
    for (int c = 0; c < n_cl; c++) free(cl_it[c]);
    free(cl_it); free(cl_size); free(tmpA); free(tmpB); free(cl_of); free(D); free(dist); free(iwn);
}

/* =====================================================================================================================================
 * ClusterOne_MyClustering.call — the clusterer of every job of MORE than 100 reads (UmiClustering.java:L240)
 *   F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class (ClusterOne_MyClustering.java:L59-L166 call, L175-L219 clusterLocal)
 *   OneUmiCluster.removeEntries / setClusterCenterNotPreGrouped (OneUmiCluster.java:L114-L119, L49-L65)
 * All of its streams are parallel above 30 reads (L176-L177, L187-L189); this restatement has the SEQUENTIAL semantics (a JVM with one worker
 * thread).  Containers whose order reaches the result: fastutil Int2ObjectOpenHashMap (possibleClusters, L185) and IntOpenHashSet (OneUmiCluster,
 * incl. java.util.AbstractCollection.removeAll driven by the set's own iterator), java.util.HashSet<Integer>, ConcurrentHashMap (idMap, L199:
 * bins in insertion order, a transfer keeps a bin's last run and prepends the nodes before it), HashSet<Set<Integer>> (L219, hash = element sum).
 * ORC_UA_TIE_UNPIN here = (a) a read had several largest neighbour sets to choose from that are not the same set (Stream.max keeps the first in the
 * map's iteration order, which a parallel toMap does not fix), or (b) a JDK/CHM bin reached the treeify threshold (not modelled). */
static uint32_t jdk_spread(uint32_t h) { return h ^ (h >> 16); }
static uint64_t sig_mix(int v)
{
    uint64_t h = (uint64_t)(v + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return h;
}
/* iteration order of a java.util.HashSet filled in the given order; hash[i] = hashCode of element i (elements are distinct) */
static void jdk_order_by_hash(const uint32_t *hash, int k, int *perm, int *long_bin)
{
    const int cap = jdk_cap_for(k);
    int *cnt = (int *)calloc((size_t)cap + 1, sizeof(int));
    for (int i = 0; i < k; i++) cnt[(jdk_spread(hash[i]) & (uint32_t)(cap - 1)) + 1]++;
    for (int b = 0; b < cap; b++) { if (cnt[b + 1] >= 9) *long_bin = 1; cnt[b + 1] += cnt[b]; }
    for (int i = 0; i < k; i++) perm[cnt[jdk_spread(hash[i]) & (uint32_t)(cap - 1)]++] = i;
    free(cnt);
}
/* iteration order of a ConcurrentHashMap<Integer, ?> filled by one thread in the given order (keys distinct) */
typedef struct { uint32_t h; int k; } chm_node;
static void chm_order(const int *keys, int K, int *out, int *long_bin)
{
    int cap = 16, sc = 12, count = 0;
    chm_node **tab = (chm_node **)calloc((size_t)cap, sizeof(chm_node *));
    int *len = (int *)calloc((size_t)cap, sizeof(int));
    for (int t = 0; t < K; t++) {
        const uint32_t h = jdk_spread((uint32_t)keys[t]) & 0x7FFFFFFFu;
        const int b = (int)(h & (uint32_t)(cap - 1));
        if (len[b] >= 8) *long_bin = 1;
        tab[b] = (chm_node *)realloc(tab[b], (size_t)(len[b] + 1) * sizeof(chm_node));
        tab[b][len[b]].h = h; tab[b][len[b]].k = keys[t]; len[b]++;
        count++;
        while (count >= sc) {                                                          /* addCount -> transfer */
            chm_node **nt = (chm_node **)calloc((size_t)cap * 2, sizeof(chm_node *));
            int *nl = (int *)calloc((size_t)cap * 2, sizeof(int));
            for (int i = 0; i < cap; i++) {
                if (!len[i]) continue;
                const chm_node *c = tab[i];
                uint32_t run_bit = c[0].h & (uint32_t)cap; int last_run = 0;
                for (int j = 1; j < len[i]; j++) { const uint32_t bb = c[j].h & (uint32_t)cap; if (bb != run_bit) { run_bit = bb; last_run = j; } }
                chm_node *lo = (chm_node *)malloc((size_t)len[i] * sizeof(chm_node)), *hi = (chm_node *)malloc((size_t)len[i] * sizeof(chm_node));
                int nlo = 0, nhi = 0;
                for (int j = last_run - 1; j >= 0; j--) {                               /* prepended one by one: they end up reversed, before the run */
                    if ((c[j].h & (uint32_t)cap) == 0) lo[nlo++] = c[j]; else hi[nhi++] = c[j];
                }
                for (int j = last_run; j < len[i]; j++) { if (run_bit == 0) lo[nlo++] = c[j]; else hi[nhi++] = c[j]; }
                nt[i] = lo; nl[i] = nlo; nt[i + cap] = hi; nl[i + cap] = nhi;
                free(tab[i]);
            }
            free(tab); free(len); tab = nt; len = nl; cap *= 2; sc = cap - (cap >> 2);
        }
    }
    int o = 0;
    for (int i = 0; i < cap; i++) { for (int j = 0; j < len[i]; j++) out[o++] = tab[i][j].k; free(tab[i]); }
    free(tab); free(len);
}

/* fastutil IntOpenHashSet with the operations OneUmiCluster sees */
typedef struct { int n, size, has_zero, min_n; int *key; } fuset;
static int fu_max_fill(int n) { const int c = (int)ceil(n * 0.75f); return c < n - 1 ? c : n - 1; }
static void fu_init(fuset *s) { s->n = 32; s->size = 0; s->has_zero = 0; s->min_n = 32; s->key = (int *)calloc(32, sizeof(int)); }
static void fu_rehash(fuset *s, int nn)
{
    int *nk = (int *)calloc((size_t)nn, sizeof(int));
    for (int j = s->n - 1; j >= 0; j--)
        if (s->key[j] != 0) { int pos = (int)(fu_mix(s->key[j]) & (uint32_t)(nn - 1)); while (nk[pos] != 0) pos = (pos + 1) & (nn - 1); nk[pos] = s->key[j]; }
    free(s->key); s->key = nk; s->n = nn;
}
static void fu_add(fuset *s, int k)
{
    if (k == 0) { if (s->has_zero) return; s->has_zero = 1; }
    else {
        int pos = (int)(fu_mix(k) & (uint32_t)(s->n - 1));
        while (s->key[pos] != 0) { if (s->key[pos] == k) return; pos = (pos + 1) & (s->n - 1); }
        s->key[pos] = k;
    }
    if (s->size++ >= fu_max_fill(s->n)) fu_rehash(s, fu_array_size(s->size + 1));
}
static int fu_order(const fuset *s, int *out)
{
    int o = 0;
    if (s->has_zero) out[o++] = 0;
    for (int j = s->n - 1; j >= 0; j--) if (s->key[j] != 0) out[o++] = s->key[j];
    return o;
}
static void fu_shift(fuset *s, int pos, int *wrapped, int *n_wrapped)
{
    const int mask = s->n - 1;
    for (;;) {
        const int last = pos;
        int curr;
        pos = (pos + 1) & mask;
        for (;;) {
            if ((curr = s->key[pos]) == 0) { s->key[last] = 0; return; }
            const int slot = (int)(fu_mix(curr) & (uint32_t)mask);
            if (last <= pos ? (last >= slot || slot > pos) : (last >= slot && slot > pos)) break;
            pos = (pos + 1) & mask;
        }
        if (wrapped && pos < last) wrapped[(*n_wrapped)++] = curr;
        s->key[last] = curr;
    }
}
static void fu_remove(fuset *s, int k)                                                  /* IntOpenHashSet.remove(int) */
{
    if (k == 0) { if (!s->has_zero) return; s->has_zero = 0; s->size--; }
    else {
        int pos = (int)(fu_mix(k) & (uint32_t)(s->n - 1));
        while (s->key[pos] != k) { if (s->key[pos] == 0) return; pos = (pos + 1) & (s->n - 1); }
        s->size--;
        fu_shift(s, pos, NULL, NULL);
    }
    if (s->n > s->min_n && s->size < fu_max_fill(s->n) / 4 && s->n > 16) fu_rehash(s, s->n / 2);
}
/* AbstractCollection.removeAll(c): walk THIS with the set's iterator, Iterator.remove() where victim[x] is set */
static void fu_remove_all(fuset *s, const uint8_t *victim)
{
    int pos = s->n, c = s->size, must_null = s->has_zero, n_wrapped = 0;
    int *wrapped = (int *)malloc((size_t)(s->size + 1) * sizeof(int));
    while (c != 0) {
        c--;
        if (must_null) { must_null = 0; if (victim[0]) { s->has_zero = 0; s->size--; } continue; }
        for (;;) {
            if (--pos < 0) { const int cur = wrapped[-pos - 1]; if (victim[cur]) fu_remove(s, cur); break; }
            if (s->key[pos] != 0) { if (victim[s->key[pos]]) { fu_shift(s, pos, wrapped, &n_wrapped); s->size--; } break; }
        }
    }
    free(wrapped);
}

/* test hook: an IntOpenHashSet filled with keys[0 .. k), its iteration order, then removeAll(victims) and the order after it.  max_key = the
 * largest key (sizes the victim map).  Returns the size after the removal. */
int orc_fu_set_ops(const int *keys, int k, const int *victims, int nv, int max_key, int *order_before, int *order_after)
{
    fuset s; fu_init(&s);
    for (int i = 0; i < k; i++) fu_add(&s, keys[i]);
    fu_order(&s, order_before);
    uint8_t *vm = (uint8_t *)calloc((size_t)max_key + 2, 1);
    for (int i = 0; i < nv; i++) vm[victims[i]] = 1;
    fu_remove_all(&s, vm);
    const int n = fu_order(&s, order_after);
    free(vm); free(s.key);
    return n;
}

typedef struct { fuset set; int center; } mycluster;

static int my_center(const int32_t *matrix, int n, const fuset *s, int qv01, int *tmp)  /* OneUmiCluster.java:L49-L65 */
{
    const int k = fu_order(s, tmp);
    if (k == 2) return qv01 ? tmp[0] : tmp[1];
    long best = -1; int center = tmp[0];
    for (int i = 0; i < k; i++) {
        long sum = 0;
        for (int j = 0; j < k; j++) if (j != i) { const int e = ed_of(matrix[(size_t)tmp[i] * n + tmp[j]]); sum += (long)e * e; }
        if (best < 0 || sum < best) { best = sum; center = tmp[i]; }
    }
    return center;
}

/* clusterLocal (L175-L219) over `idx` (L indices in the caller's order).  Clusters come back as one flat member array (each in the iteration order of
 * its HashSet<Integer>) with n_cl + 1 offsets, in the iteration order of the HashSet<Set<Integer>>.  Returns n_cl. */
static int my_cluster_local(const int32_t *matrix, int n, const int *idx, int L, int ed, int *members, int *cl_off, int *harmful, int *long_bin)
{
    int *cnt = (int *)calloc((size_t)n, sizeof(int)), *keys_in = (int *)malloc((size_t)L * sizeof(int)), nk = 0;
    uint64_t *sig = (uint64_t *)calloc((size_t)n, sizeof(uint64_t));
    /* (a giant job is handed over outside the batch's parallel loop, orc_umi_assign_batch: its two O(L^2) passes then use all threads) */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (L > 1500)
#endif
    for (int i = 0; i < L; i++) {                                                       /* L179-L184 */
        const int a = idx[i];
        int c = 0; uint64_t sg = 0;
        for (int j = 0; j < L; j++) if (ed_of(matrix[(size_t)a * n + idx[j]]) <= ed) { c++; sg += sig_mix(idx[j]); }
        cnt[a] = c; sig[a] = sg;
    }
    for (int i = 0; i < L; i++) if (cnt[idx[i]] > 1) keys_in[nk++] = idx[i];
    if (nk == 0) { free(cnt); free(keys_in); free(sig); return 0; }
    int *keys = (int *)malloc((size_t)nk * sizeof(int));
    {                                                                                   /* Int2ObjectOpenHashMap: same layout and iteration as the set */
        fuset m; fu_init(&m);
        for (int i = 0; i < nk; i++) fu_add(&m, keys_in[i]);
        fu_order(&m, keys); free(m.key);
    }
    int *chosen = (int *)malloc((size_t)nk * sizeof(int));
    int harm = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(|:harm) if (nk > 1500)
#endif
    for (int i = 0; i < nk; i++) {                                                      /* L190-L196 */
        const int c = keys[i];
        int best = -1;
        for (int j = 0; j < nk; j++) { const int e = keys[j]; if (ed_of(matrix[(size_t)e * n + c]) <= ed && (best < 0 || cnt[e] > cnt[best])) best = e; }
        for (int j = 0; j < nk; j++) { const int e = keys[j]; if (ed_of(matrix[(size_t)e * n + c]) <= ed && cnt[e] == cnt[best] && sig[e] != sig[best]) harm |= 1; }
        chosen[i] = best;
    }
    if (harm) *harmful = 1;
    /* L199: groups in first-seen order, members in key order */
    int *gid = (int *)malloc((size_t)n * sizeof(int)), *first = (int *)malloc((size_t)nk * sizeof(int)), ng = 0;
    for (int i = 0; i < n; i++) gid[i] = -1;
    for (int i = 0; i < nk; i++) if (gid[chosen[i]] < 0) { gid[chosen[i]] = ng; first[ng++] = chosen[i]; }
    int *gsz = (int *)calloc((size_t)ng + 1, sizeof(int));
    for (int i = 0; i < nk; i++) gsz[gid[chosen[i]] + 1]++;
    for (int g = 0; g < ng; g++) gsz[g + 1] += gsz[g];
    int *gm = (int *)malloc((size_t)nk * sizeof(int)), *fill = (int *)malloc((size_t)ng * sizeof(int));
    for (int g = 0; g < ng; g++) fill[g] = gsz[g];
    for (int i = 0; i < nk; i++) gm[fill[gid[chosen[i]]]++] = keys[i];
    int *corder = (int *)malloc((size_t)ng * sizeof(int));
    chm_order(first, ng, corder, long_bin);                                             /* idMap.values() */
    uint32_t *hs = (uint32_t *)calloc((size_t)ng + 1, sizeof(uint32_t));
    for (int t = 0; t < ng; t++) { const int g = gid[corder[t]]; uint32_t h = 0; for (int i = gsz[g]; i < gsz[g + 1]; i++) h += (uint32_t)gm[i]; hs[t] = h; }
    int *perm = (int *)malloc((size_t)ng * sizeof(int));
    jdk_order_by_hash(hs, ng, perm, long_bin);                                          /* L219: collect(toSet()) of the value sets */
    int o = 0;
    cl_off[0] = 0;
    for (int t = 0; t < ng; t++) {
        const int g = gid[corder[perm[t]]], k = gsz[g + 1] - gsz[g];
        uint32_t *hh = (uint32_t *)malloc((size_t)k * sizeof(uint32_t)); int *pp = (int *)malloc((size_t)k * sizeof(int));
        for (int i = 0; i < k; i++) hh[i] = (uint32_t)gm[gsz[g] + i];
        jdk_order_by_hash(hh, k, pp, long_bin);                                         /* mapping(left, toSet()): HashSet<Integer> in key order */
        for (int i = 0; i < k; i++) members[o++] = gm[gsz[g] + pp[i]];
        cl_off[t + 1] = o;
        free(hh); free(pp);
    }
    free(cnt); free(keys_in); free(sig); free(keys); free(chosen); free(gid); free(first); free(gsz); free(gm); free(fill); free(corder); free(hs); free(perm);
    return ng;
}

void orc_umi_assign_myclust(const int32_t *matrix, int64_t n64, const orc_assign_params *P, int qv01, orc_assign_rec *rec)
{
    const int n = (int)n64, ed = P->ed_complete;                                        /* ctor L52 */
    for (int i = 0; i < n; i++) { memset(&rec[i], 0, sizeof(rec[i])); rec[i].center = -1; rec[i].u2 = -1; rec[i].flags = ORC_UA_DEEP; }
    if (n < 1) return;
    int harmful = 0, long_bin = 0;
    int *idx = (int *)malloc((size_t)n * sizeof(int)), *mem = (int *)malloc((size_t)n * sizeof(int)), *off = (int *)malloc((size_t)(n + 1) * sizeof(int));
    int *tmp = (int *)malloc((size_t)(n + 1) * sizeof(int));
    for (int i = 0; i < n; i++) idx[i] = i;
    const int n_full = my_cluster_local(matrix, n, idx, n, ed, mem, off, &harmful, &long_bin);       /* L72-L73 */
    mycluster *cl = (mycluster *)malloc((size_t)(2 * n + 1) * sizeof(mycluster));
    int n_cl = 0;
    if (n_full > 0) {
        int maxdepth = 0;
        for (int c = 0; c < n_full; c++) if (off[c + 1] - off[c] > maxdepth) maxdepth = off[c + 1] - off[c];       /* L77 */
        uint8_t *clustered = (uint8_t *)calloc((size_t)n, 1), *victim = (uint8_t *)calloc((size_t)n, 1);
        for (int c = 0; c < n_full; c++) {                                              /* L78-L88 */
            const int k = off[c + 1] - off[c];
            if (!((long)k * P->fold_depth > maxdepth)) {
                for (int i = off[c]; i < off[c + 1]; i++) { rec[mem[i]].flags |= ORC_UA_SKIPPED; rec[mem[i]].cluster_size = (uint16_t)(k > 65535 ? 65535 : k); }
                continue;
            }
            fu_init(&cl[n_cl].set);
            for (int i = off[c]; i < off[c + 1]; i++) { fu_add(&cl[n_cl].set, mem[i]); clustered[mem[i]] = 1; }
            cl[n_cl].center = my_center(matrix, n, &cl[n_cl].set, qv01, tmp);
            n_cl++;
        }
        int nu = 0, n_removed = 0;
        for (int d = 0; d < n; d++) if (!clustered[d]) idx[nu++] = d;                   /* L91 */
        for (int c = 0; c < n_cl; c++) {                                                /* L102, L60-L65 */
            const int k = fu_order(&cl[c].set, tmp);
            int nr = 0;
            for (int i = 0; i < k; i++) if (ed_of(matrix[(size_t)tmp[i] * n + cl[c].center]) > ed) { victim[tmp[i]] = 1; idx[nu++] = tmp[i]; nr++; }
            if (nr) {
                fu_remove_all(&cl[c].set, victim);
                for (int i = 0; i < k; i++) victim[tmp[i]] = 0;
                cl[c].center = my_center(matrix, n, &cl[c].set, qv01, tmp);
                n_removed += nr;
            }
        }
        if (n_removed > 0) {                                                            /* L106-L112 */
            const int n_extra = my_cluster_local(matrix, n, idx, nu, ed, mem, off, &harmful, &long_bin);
            for (int c = 0; c < n_extra; c++) {
                if (off[c + 1] - off[c] <= 1) continue;                                 /* L109 */
                fu_init(&cl[n_cl].set);
                for (int i = off[c]; i < off[c + 1]; i++) fu_add(&cl[n_cl].set, mem[i]);
                cl[n_cl].center = my_center(matrix, n, &cl[n_cl].set, qv01, tmp);
                n_cl++;
            }
        }
        uint8_t *inside = clustered;
        for (int c = 0; c < n_cl; c++) {                                                /* L116-L164 */
            const int k = fu_order(&cl[c].set, tmp), center = cl[c].center;
            if (k <= 1) continue;                                                       /* L117 (userObject is absent: no pre-grouping) */
            long sum = 0, cntv = 0;
            for (int i = 0; i < k; i++) if (tmp[i] != center) { sum += pos1_offset(matrix[(size_t)center * n + tmp[i]]); cntv++; }
            const int off_mean = (int)floor((double)sum / (double)cntv + 0.5);          /* L126-L130 */
            int nf = 0;
            for (int i = 0; i < k; i++) if (ed_of(matrix[(size_t)tmp[i] * n + center]) <= ed) nf++;                 /* L135 */
            if (nf <= 1) continue;                                                      /* L139 */
            memset(inside, 0, (size_t)n);
            for (int i = 0; i < k; i++) inside[tmp[i]] = 1;
            for (int i = 0; i < k; i++) {
                const int x = tmp[i];
                if (ed_of(matrix[(size_t)x * n + center]) > ed) continue;
                orc_assign_rec *r = &rec[x];
                if (r->flags & ORC_UA_SKIPPED) continue;                                /* ClusterOneBase.java:L122-L123 */
                r->center = center; r->flags |= ORC_UA_ASSIGNED; r->cluster_size = (uint16_t)(k > 65535 ? 65535 : k); r->off_mean = (int8_t)off_mean;
                r->u1 = (int8_t)ed_of(matrix[(size_t)center * n + x]);
                r->pos2 = (int8_t)pos2_code(matrix[(size_t)center * n + x]);
                if (n_cl > 1) {                                                         /* L161-L164 */
                    int best = -1;
                    for (int y = 0; y < n; y++) if (!inside[y]) { const int e = ed_of(matrix[(size_t)x * n + y]); if (best < 0 || e < best) best = e; }
                    r->u2 = (int8_t)best;
                }
            }
        }
        free(clustered); free(victim);
    }
    for (int i = 0; i < n; i++) { rec[i].n_clusters = n_cl; if (harmful || long_bin) rec[i].flags |= ORC_UA_TIE_UNPIN; }
    for (int c = 0; c < n_cl; c++) free(cl[c].set.key);
    free(cl); free(idx); free(mem); free(off); free(tmp);
}

void orc_umi_assign_batch(const int32_t *matrices, const int64_t *job_offsets, const int64_t *out_offsets, int64_t n_jobs,
                          const orc_assign_params *P, const uint8_t *job_qv01, orc_assign_rec *rec, int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int64_t j = 0; j < n_jobs; j++) {
        const int64_t r0 = job_offsets[j], n = job_offsets[j + 1] - r0;
        if (n > P->max_hier && P->deep) {                                                /* UmiClustering.java:L240: ClusterOne_MyClustering's job */
            if (n <= 1500) orc_umi_assign_myclust(matrices + out_offsets[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0);
            continue;                                                                    /* giant jobs: below, one at a time with all threads inside */
        }
        if (n > P->max_hier) {
            for (int64_t i = 0; i < n; i++) { memset(&rec[r0 + i], 0, sizeof(rec[0])); rec[r0 + i].center = -1; rec[r0 + i].u2 = -1; rec[r0 + i].flags = ORC_UA_DEEP; }
            continue;
        }
        orc_umi_assign_hier(matrices + out_offsets[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0);
    }
    if (P->deep)
        for (int64_t j = 0; j < n_jobs; j++) {
            const int64_t r0 = job_offsets[j], n = job_offsets[j + 1] - r0;
            if (n > P->max_hier && n > 1500) orc_umi_assign_myclust(matrices + out_offsets[j], n, P, job_qv01 ? job_qv01[j] : 0, rec + r0);
        }
}
