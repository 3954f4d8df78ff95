/*
 * sicelore_gpu.h — C ABI of libsicelore_gpu.so, the B200 (sm_100a) drop-in for the barcode / UMI
 * edit-distance hot path of SiCeLoRe 2.1's NanoporeBC_UMI_finder (scanfastq pass 2, assignumis).
 *
 * The reference has no FFI for this path (pure in-process Java); each entry point below replaces one
 * loop body of the reference and is what a JNI `native` method of a `com.rw.gpu.Native` class binds
 * (see INTEGRATION.md).  Citations: F! = Jar/NanoporeBC_UMI_finder-2.1.jar, (File.java:Lnnn) = original
 * source lines recovered from the class files' LineNumberTable.
 *
 * Conventions: plain pointers and sizes; the caller owns every buffer; nothing is retained after a call
 * returns except objects behind the opaque handles; every function returns 0 on success or a negative
 * SLR_E_* code, with a thread-local message available from slr_last_error().  There is NO CPU fallback:
 * without a usable CUDA device every compute call fails with SLR_E_NODEVICE.
 * Thread safety: all entry points may be called concurrently; calls that share one context are
 * serialised internally per stream slot (see slr_ctx_create `n_streams`).
 */
#ifndef SICELORE_GPU_H
#define SICELORE_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLR_ABI_VERSION 1

enum {
    SLR_OK = 0,
    SLR_E_INVALID = -1,      /* bad argument */
    SLR_E_NODEVICE = -2,     /* no CUDA device / driver: the library never falls back to the CPU */
    SLR_E_CUDA = -3,         /* a CUDA runtime call failed (message has the detail) */
    SLR_E_NOMEM = -4,
    SLR_E_UNSUPPORTED = -5   /* e.g. --bcEditDistance > 2, barcode length != 16 */
};

typedef struct slr_ctx slr_ctx;           /* one per (process, device) */
typedef struct slr_bc_table slr_bc_table; /* device-resident search set (whitelist or used-barcode list) */

/* ---- per-read result of Parser.assignBarcode (F!…/analyzers/Parser.class, Parser.java:L244-L311) ------ */
#define SLR_F_ASSIGNED  1u   /* BC_FOUND: best.ED <= bcEditDistance && (no second || best.ED < second.ED)  (L251-L252) */
#define SLR_F_EXCEPTION 2u   /* the Java would have thrown for this read (slice too short / non-IUPAC char) */
#define SLR_F_TIE_UNPIN 4u   /* >= 11 same-hash OneMatch entries: JDK HashMap treeified its bin, tie order not emulated */

typedef struct {
    uint64_t bc;          /* OneMatch.matchingBC of the best match (2-bit packed, A=0 G=1 C=2 T=3, first base most significant); 0 if unassigned */
    int32_t  ed;          /* best edit distance over all offsets; -1 = nothing matched */
    int32_t  ed_second;   /* best ED of a DIFFERENT barcode (read-name field ed_sec=), INT32_MAX = none (L288-L289) */
    int8_t   offset;      /* OneMatch.offsetFromPredicted of the best match (L275-L276) */
    int8_t   n_ins;       /* OneMatch.insertions  \  getOffsetForReadEnd() = n_ins - n_del  (BarcodeMatchTester.java:L533) */
    int8_t   n_del;       /* OneMatch.deletions   /  bcEnd = bcStart -/+ 15 -/+ (n_ins - n_del)  (Parser.java:L278-L279) */
    int8_t   n_sub;
    int32_t  rank;        /* CountsRank.rank of bc (L267-L269); -1 if unassigned */
    uint32_t flags;       /* SLR_F_* */
} slr_bc_result;          /* 32 bytes */

/* ---- context ---------------------------------------------------------------------------------------- */

/* Replaces nothing in the reference (there is no device there); the JNI shim calls it once from
 * WorkerReadscanner's constructor (F!…/WorkerReadscanner.class, WorkerReadscanner.java:L186-L190), where the
 * reference creates its two work-stealing pools.  device < 0 selects the current device.  n_streams = how
 * many host threads may have a batch in flight at once (the reference's nCPU Parser workers).
 * The library contains sm_100a code only: a device of any other compute capability is refused with SLR_E_UNSUPPORTED.
 * Tables, candidate sets and sessions belong to the context they were created with; passing them to another context is SLR_E_INVALID. */
int  slr_ctx_create(int device, int n_streams, slr_ctx **out);
void slr_ctx_destroy(slr_ctx *ctx);
int  slr_ctx_device(const slr_ctx *ctx);

/* ---- search set ------------------------------------------------------------------------------------- */

/* Replaces BarcodesMapForBCfinding (the Long2ObjectOpenHashMap<CountsRank> built at WorkerReadscanner.java:L264-L269
 * or by getMapFromCellRangerData for --cellRangerBCs) whose keySet() is the searchSet of every
 * BarcodeMatchTester (Parser.java:L228).  barcodes2bit[i] is the reference's 2-bit `long`; rank[i] may be NULL.
 * Duplicates keep the first index.  bc_len must be 16 (config.xml:189 cell_bc_length). */
int  slr_bc_table_create(slr_ctx *ctx, const uint64_t *barcodes2bit, const int32_t *rank, int64_t n, int bc_len,
                         slr_bc_table **out);
void slr_bc_table_destroy(slr_bc_table *t);
int64_t slr_bc_table_size(const slr_bc_table *t);

/* ---- S1: barcode assignment ------------------------------------------------------------------------- */

/* Replaces the body of Parser.assignBarcode from window extraction to the best / second-best decision
 * (Parser.java:L198-L252) for a whole ReadChunk at once, i.e. 2*plusminus+1 BarcodeMatchTester.doJob runs
 * per read (BarcodeMatchTester.java:L198-L244) plus the Matches merge.
 *   ed_max      --bcEditDistance / assignCellBCwithEditDistance (0, 1 or 2)
 *   plusminus   config.xml:35 testPlusMinusPos (0..4; reference default 2)
 *   three_prime scantype == THREEP_BARCODE (Parser.java:L205)
 *   slices      n * stride bytes, ASCII, a piece of getStrandedSeq() per read
 *   slice_len   valid bytes per slice (<= stride); lens (nullable) overrides it per read
 *   anchor[i]   0-based index inside slice i of the first base of the offset-0 window:
 *               3': (adapterpos - 16) - 1 - slice_start,  5': adapterpos - slice_start   (1-based adapterpos, L206-L210)
 *               needed span: 3' [anchor-plusminus-4, anchor+plusminus+16)   5' [anchor-plusminus, anchor+plusminus+21)
 *   out         n records, positional
 * Host pointers; H2D / D2H copies are part of the call. */
int  slr_bc_assign(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime,
                   const uint8_t *slices, int stride, int slice_len, const int32_t *lens, const int32_t *anchor,
                   int64_t n, slr_bc_result *out);

/* Same, all pointers are DEVICE pointers on ctx's device, asynchronous on `stream` (a cudaStream_t cast to
 * void*, NULL = default stream).  For callers that already keep reads in HBM (bench `value`, multi-batch pipelines). */
int  slr_bc_assign_dev(slr_ctx *ctx, const slr_bc_table *t, int ed_max, int plusminus, int three_prime,
                       const uint8_t *d_slices, int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor,
                       int64_t n, slr_bc_result *d_out, void *stream);

/* Replaces the assignedBarcodes2ndPass ConcurrentHashMap<Long, BarcodeCounts> updates (Parser.java:L305-L311,
 * BarcodeCounts.addCountForEd) that feed BarcodesAssigned.tsv: the table accumulates, on the device, one
 * counter per (barcode index, ED 0..2) for every assigned read of every slr_bc_assign* call.
 * counts_out: n_barcodes * 3 int64 (host).  slr_bc_counts_device returns the device buffer itself so that
 * a multi-GPU driver can all-reduce it (NCCL) before reading it back.
 * slr_bc_counts_read is ordered behind every host-pointer call issued on this context before it (no device-wide synchronisation);
 * launches of the *_dev entry points run on the caller's streams, which the caller synchronises first. */
int  slr_bc_counts_read(slr_ctx *ctx, const slr_bc_table *t, int64_t *counts_out);
int  slr_bc_counts_reset(slr_ctx *ctx, slr_bc_table *t);
int  slr_bc_counts_device(const slr_bc_table *t, int64_t **d_counts, int64_t *n_elems);

/* ---- pass 1: exact lookup of the offset-0 window (used-barcode counting) ---------------------------------- */

/* Replaces the per-read body of UsedCellBCListGenerator$Worker.call (F!com/rw/nanoporereadscanner/analyzers/
 * UsedCellBCListGenerator$Worker.class, UsedCellBCListGenerator.java:L206-L232): the window at the predicted position
 * (3': read[adapterpos-16 .. adapterpos-1] reverse-complemented, 5': read[adapterpos+1 .. adapterpos+16]) is looked up
 * in the 10x whitelist and counted.  Same buffers as slr_bc_assign (anchor = window start in the slice); no post
 * sequence is taken, so no flank is needed.  out[i].flags & SLR_F_ASSIGNED = in the list (bc, rank set, ed = 0); the
 * per-barcode counts (unfilteredUsedBarcodeMap) accumulate in the table's ED-0 counters (slr_bc_counts_read). */
int  slr_bc_exact(slr_ctx *ctx, const slr_bc_table *t, int three_prime, const uint8_t *slices, int stride, int slice_len,
                  const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out);
int  slr_bc_exact_dev(slr_ctx *ctx, const slr_bc_table *t, int three_prime, const uint8_t *d_slices, int stride, int slice_len,
                      const int32_t *d_lens, const int32_t *d_anchor, int64_t n, slr_bc_result *d_out, void *stream);

/* ---- S3: pass-1 collision test of the used-barcode list ------------------------------------------------- */

/* Matches of one BarcodeMatchTester run of the collision tester: at most one OneMatch per ED level
 * (BarcodeMatchTester$Matches is a HashSet whose equals() is (readSeq, ED, offset), BarcodeMatchTester.java:L433-L436). */
typedef struct {
    uint64_t bc[2];      /* OneMatch.matchingBC of the ED-1 / ED-2 entry */
    uint8_t  valid;      /* bit 0: ED-1 entry present, bit 1: ED-2 entry present */
    uint8_t  n_sub[2], n_ins[2], n_del[2];   /* OneMatch counters as the Java names them */
    uint8_t  pad;
} slr_collide_result;    /* 24 bytes */

/* Replaces the BarcodeDatasetColissionTester.submitSeq loop (F!com/rw/nanoporereadscanner/analyzers/
 * BarcodeDatasetColissionTester.class, BarcodeDatasetColissionTester.java:L212-L229): for every barcode of the
 * used-barcode list one BarcodeMatchTester(seq, editDistance, skipFullMatches=true, allowIndels=true,
 * searchSet = the list's keySet(), offset 0, cell_bc_length, postSeq=null, doNextLevelIfMatchFound=false).call()
 * (L215-L222), whose Matches feed getUnfilteredColissionData / generateColissionMergedBCmap (L126-L203, host Java).
 *   t          the search set = barcodes_b4filtering.keySet()  (slr_bc_table_create of the same list)
 *   ed_max     mergeBCsED (config.xml:25; null = --bcEditDistance): 0, 1 or 2
 *   barcodes   n queries (2-bit longs), normally the list itself; out: n records, positional */
int  slr_bc_collide(slr_ctx *ctx, const slr_bc_table *t, int ed_max, const uint64_t *barcodes, int64_t n,
                    slr_collide_result *out);
int  slr_bc_collide_dev(slr_ctx *ctx, const slr_bc_table *t, int ed_max, const uint64_t *d_barcodes, int64_t n,
                        slr_collide_result *d_out, void *stream);

/* ---- S2: UMI distance matrices ---------------------------------------------------------------------- */

/* Replaces ClusteringEditDistanceBase.generateDistanceMatrix (F!com/rw/clustering/ClusteringEditDistanceBase.class,
 * ClusteringEditDistanceBase.java:L168-L259) for ALL (cell, region) jobs of one BAM chunk
 * (UmiClustering.cluster, UmiClustering.java:L131-L146): per read pair the 3x3 shifted thresholded
 * Levenshtein distances (calcEditDistances, L297-L350; limitedCompare threshold 4, -1 -> 5) reduced to the
 * packed BestEditDistance int (L67-L80, L425-L428).
 *   umis        m * stride bytes; per read umi_len+2 4-bit codes (A=1 G=2 C=4 T=8 N=15 ...,
 *               NucleicAcidByteCodeBase.java:L45-L78) = getSubSequence(bcEnd, umi_len+2) of the strand-corrected
 *               X= mini sequence, i.e. the predicted UMI window widened by one base on each side
 *   umi_len     config.xml:264 umi_length, 1..14 (the nine comparisons of a pair run on one 32-bit word; longer UMIs: SLR_E_UNSUPPORTED)
 *   job_offsets n_jobs+1 CSR offsets into the reads
 *   out_offsets n_jobs+1 offsets into `out`; job j writes an n_j x n_j row-major int32 matrix at out_offsets[j]
 *               (diagonal = the reference's equalityEditDistance, lower triangle = transposed copy) */
int  slr_umi_dist(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets,
                  int64_t n_jobs, int32_t *out, const int64_t *out_offsets);
int  slr_umi_dist_dev(slr_ctx *ctx, const uint8_t *d_umis, int stride, int umi_len, const int64_t *d_job_offsets,
                      int64_t n_jobs, int64_t n_reads, int32_t *d_out, const int64_t *d_out_offsets,
                      int64_t n_out, void *stream);

/* ---- S5: neighbour-set clustering on the matrices ------------------------------------------------- */

/* Replaces the two O(n^2) steps of ClusterOne_MyClustering.clusterLocal (F!com/rw/umifinder/analyzers/clustering/
 * ClusterOne_MyClustering.class, ClusterOne_MyClustering.java:L175-L219) for ALL jobs of one BAM chunk, fused behind
 * the S2 matrices (which stay on the device):
 *   possibleClusters  a -> N(a) = { v in indices : getED(matrix[a][v]) <= ed }, kept when |N(a)| > 1   (L179-L185;
 *                     getED = (byte)(packed & 0xFFFFFF), ClusteringEditDistanceBase$BestEditDistance.java:L382)
 *   per key c         the entry l with c in N(l) and the largest |N(l)|; Stream.max keeps the FIRST maximum of
 *                     possibleClusters.int2ObjectEntrySet() (L190-L196)
 * The grouping of the keys by that entry (L199, L219) is an O(n) pass that stays with the caller.
 *   umis, stride, umi_len, job_offsets, n_jobs   as slr_umi_dist
 *   ed          ClusterOne_MyClustering.ed (config.xml:270-272: 2 = first pass, 1 = second), 0..5
 *   member      NULL = every read is in `indices`; else m bytes, 1 = in `indices` (the re-clustering call of
 *               ClusterOne_MyClustering.call, …java:L107, passes the unclustered subset)
 *   rank        NULL, or m int32: iteration rank of key l in the caller's possibleClusters map (smaller = earlier;
 *               job-local).  The fastutil slot order depends on how the map was filled (for > 30 reads the
 *               reference fills it from a parallel stream), so only the caller knows it.  With NULL a tie goes to
 *               the smallest index and n_ties > 1 marks exactly the reads whose choice depends on that order.
 *   out, out_offsets   the matrices as slr_umi_dist writes them, or out = NULL to leave them on the device
 *   rec         m records, positional */
typedef struct slr_umi_cluster_rec {
    int32_t n_neighbours;    /* |N(c)|, the read itself included when matrix[c][c] <= ed; 0 for a non-member        */
    int32_t best_key;        /* job-local index of the chosen entry; -1 when c is no key (|N(c)| <= 1)             */
    int32_t best_count;      /* |N(best_key)|                                                                       */
    int32_t n_ties;          /* entries containing c whose |N| equals best_count                                    */
} slr_umi_cluster_rec;       /* 16 bytes */
int  slr_umi_cluster(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets,
                     int64_t n_jobs, int ed, const uint8_t *member, const int32_t *rank, int32_t *out,
                     const int64_t *out_offsets, slr_umi_cluster_rec *rec);
/* the same two steps on matrices already on the device (as slr_umi_dist_dev left them); d_counts: m int32 scratch */
int  slr_umi_cluster_dev(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets,
                         const int64_t *d_out_offsets, int64_t n_jobs, int64_t n_reads, int ed,
                         const uint8_t *d_member, const int32_t *d_rank, int32_t *d_counts,
                         slr_umi_cluster_rec *d_rec, void *stream);

/* A session keeps the matrices of one batch of jobs on the device between calls.  clusterLocal's first maximum depends on the
 * iteration order of the caller's map, ties are frequent (two reads that are each other's only neighbour already tie), and the
 * keys of that map are known only after the neighbour counts: the practical protocol is  create  ->  cluster(rank = NULL)  ->
 * the caller fills its map with the keys (n_neighbours > 1) and reads their iteration ranks  ->  cluster(rank)  [-> cluster(member,
 * rank) for the re-clustering call]  ->  matrices (when the distances are needed on the host)  ->  destroy, with the distance kernels run once.
 *   create     umis / stride / umi_len / job_offsets / n_jobs as slr_umi_dist; uploads the reads and computes every matrix
 *   cluster    as slr_umi_cluster on the resident matrices; callable any number of times
 *   matrices   copies the matrices back, packed back to back in job order (job j at sum of n_k^2 over k < j)
 * Device memory: 4 bytes per matrix cell + 61 bytes per read; a batch that does not fit fails with SLR_E_NOMEM (split it). */
typedef struct slr_umi_session slr_umi_session;
int  slr_umi_session_create(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets,
                            int64_t n_jobs, slr_umi_session **out);
int  slr_umi_session_cluster(slr_umi_session *s, int ed, const uint8_t *member, const int32_t *rank,
                             slr_umi_cluster_rec *rec);
int  slr_umi_session_matrices(slr_umi_session *s, int32_t *out, int64_t n_cells);
int64_t slr_umi_session_cells(const slr_umi_session *s);
int64_t slr_umi_session_reads(const slr_umi_session *s);   /* reads / jobs of the batch: the sizes the cluster / assign record buffers need */
int64_t slr_umi_session_jobs(const slr_umi_session *s);
void slr_umi_session_destroy(slr_umi_session *s);

/* ---- S6: clustering of the small jobs + UMI assignment --------------------------------------------------------- */

/* Replaces ClusterOneHierarchical.call (F!com/rw/umifinder/analyzers/clustering/ClusterOneHierarchical.class,
 * ClusterOneHierarchical.java:L61-L217) — the clusterer UmiClustering$Submitter picks for every job of at most 100 reads
 * (UmiClustering.java:L239-L261; its static CLUSTERHOW is DECIDEONCOMPLEXITY, L50, so the pre-grouping branch L242-L244 is unreachable) —
 * for ALL such jobs of a BAM chunk, fused behind the S2 matrices:
 *   reads with a neighbour within umi_completelinkclusteringED              DistanceMatrix.java:L87-L90
 *   LingPipe CompleteLinkClusterer (or SingleLinkClusterer above the switch threshold) on them, Dendrogram.partitionDistance(ED), clusters of
 *   more than one read                                                      A!com/aliasi/cluster/CompleteLinkClusterer.java:L146-L237, Dendrogram.java:L205-L215
 *   depth rule: size * foldDepthBelowMaxDiscardForClustering > largest cluster, else flagDontUMIassignRecords   ClusterOneHierarchical.java:L118-L127
 *   OneUmiCluster.setClusterCenter                                          F!com/rw/clustering/OneUmiCluster.java:L49-L65
 *   per read what ClusterOneBase.setSamflagsAndStatsForClustered derives    ClusterOneBase.java:L118-L168
 * The caller keeps: the strings (U8 = getPostBCUMIseqOffset(centre read, offset_center_mean), U7 = the read's own window), the noUMIsoFar /
 * SKIPPED_HIGHCOMPLEXITY guards (ClusterOneBase.java:L118-L123) and the statistics counters. */
#define SLR_UA_ASSIGNED  1u   /* the read is in a cluster of the final list: setSamflagsAndStatsForClustered runs for it */
#define SLR_UA_SKIPPED   2u   /* its cluster failed the depth rule: UMI_CLUSTERING_SKIPPED_HIGHCOMPLEXITY is set on the read */
#define SLR_UA_TIE_UNPIN 4u   /* the reference's own result for this job depends on JVM identity hash codes (LingPipe's ObjectToSet keeps its
                                 PairScores in a HashSet without hashCode()): equal-cost pairs created by one merge have no defined queue order
                                 there.  The records follow creation order; a caller that wants its JVM's choice re-runs exactly these jobs. */
#define SLR_UA_DEEP      8u   /* job of more than max_hier reads: ClusterOne_MyClustering's.  With params.deep (the default) its records are
                                 filled by the large-job path below and carry this bit too; with deep = 0 (or no arena, slr_umi_assign_dev)
                                 the job is only flagged and left to slr_umi_cluster / the session API */
typedef struct slr_umi_assign_params {
    int32_t ed_complete;      /* umi_completelinkclusteringED (config.xml:270), also the neighbour threshold */
    int32_t ed_single;        /* umi_singlelinkclusteringED (config.xml:272) */
    int32_t single_threshold; /* complexity_threshold_for_switch_to_single_link_clustering (config.xml:278: 3000, i.e. never for <= 100 reads) */
    int32_t fold_depth;       /* foldDepthBelowMaxDiscardForClustering (UMIparameters.java:L118: 50) */
    int32_t max_hier;         /* jobs up to this size are ClusterOneHierarchical's (UmiClustering.java:L240: 100; at most 100 here) */
    int32_t deep;             /* 1: jobs above max_hier run ClusterOne_MyClustering.call on the GPU (see below); 0: they are only flagged */
} slr_umi_assign_params;      /* NULL = { 2, 1, 3000, 50, 100, 1 } */
/* Jobs above max_hier: ClusterOne_MyClustering.call (F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class,
 * ClusterOne_MyClustering.java:L59-L166) as a whole — clusterLocal over all reads (L175-L219), the depth rule (L77-L84), setClusterCenter, the
 * off-centre removal (L60-L65, OneUmiCluster.removeEntries L114-L119), clusterLocal over the unclustered reads (L104-L112) and the per-read
 * values (L116-L164).  Every stream of that class is parallel above 30 reads, so the reference's own result depends on thread timing wherever
 * an iteration order decides; the records have the SEQUENTIAL semantics (a JVM with one worker thread), with the orders of fastutil's
 * Int2ObjectOpenHashMap / IntOpenHashSet (incl. iterator-driven removeAll), java.util.HashSet and ConcurrentHashMap reproduced.  SLR_UA_TIE_UNPIN
 * on such a job = a read could choose between largest neighbour sets that are not the same set (Stream.max keeps the first in map order),
 * or a hash bin reached the JDK's treeify threshold (not modelled). */
typedef struct slr_umi_assign_rec {
    int32_t  center;              /* job-local index of OneUmiCluster.getCenter() of the read's cluster (U8 comes from that read); -1 = none */
    int8_t   u1;                  /* UMI_ED (U1): distanceNonReducedSet(center, read) */
    int8_t   u2;                  /* UMI_ED_SECOND_BEST_MATCH (U2): least distance to a read outside the cluster; -1 = the tag is not written */
    int8_t   pos2;                /* matrix[center][read].getPos2(): 0 MINUSONE, 1 ZERO, 2 PLUSONE (the PREDICTED_POS statistics flag) */
    int8_t   offset_center_mean;  /* offsetcentermean of the cluster: round(mean getPos1().getOffSet() of matrix[center][member]) */
    uint16_t flags;               /* SLR_UA_* */
    uint16_t cluster_size;
    int32_t  n_clusters;          /* cluster_list.size() of the job (U2 is written only when > 1) */
} slr_umi_assign_rec;             /* 16 bytes */
/*   umis, stride, umi_len, job_offsets, n_jobs   as slr_umi_dist
 *   job_qv01   NULL or one byte per job: mean_qv(job's read 0) > mean_qv(job's read 1) — the rule OneUmiCluster.java:L53 uses for clusters of two
 *   out, out_offsets   the matrices as slr_umi_dist writes them, or out = NULL to leave them on the device
 *   rec        one record per read, positional */
int  slr_umi_assign(slr_ctx *ctx, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                    const slr_umi_assign_params *params, const uint8_t *job_qv01, int32_t *out, const int64_t *out_offsets,
                    slr_umi_assign_rec *rec);
/* the same on matrices already on the device (as slr_umi_dist_dev left them); d_scratch: slr_umi_assign_scratch_bytes(n_jobs) bytes.
 * slr_umi_assign_dev has no room for the working arrays of the jobs above max_hier: they are only flagged SLR_UA_DEEP.  slr_umi_assign_dev2 takes
 * scratch_bytes = slr_umi_assign_scratch_bytes(n_jobs) + the sum of slr_umi_assign_deep_job_bytes(n) over the jobs above max_hier (any upper
 * bound will do; a deep job that does not fit is left flagged, n_clusters 0, never half-written).  slr_umi_assign_deep_job_bytes(n) is about
 * 300 n + n * n / 8 bytes: lists and hash tables of the emulated containers + one threshold bit per matrix cell. */
int64_t slr_umi_assign_scratch_bytes(int64_t n_jobs);
int64_t slr_umi_assign_deep_job_bytes(int64_t n_reads_of_job);
int  slr_umi_assign_dev(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets, const int64_t *d_out_offsets, int64_t n_jobs,
                        int64_t n_reads, const slr_umi_assign_params *params, const uint8_t *d_job_qv01, void *d_scratch,
                        slr_umi_assign_rec *d_rec, void *stream);
int  slr_umi_assign_dev2(slr_ctx *ctx, const int32_t *d_matrices, const int64_t *d_job_offsets, const int64_t *d_out_offsets, int64_t n_jobs,
                         int64_t n_reads, const slr_umi_assign_params *params, const uint8_t *d_job_qv01, void *d_scratch, int64_t scratch_bytes,
                         slr_umi_assign_rec *d_rec, void *stream);
/* on the resident matrices of a session (see below) */
struct slr_umi_session;
int  slr_umi_session_assign(struct slr_umi_session *s, const slr_umi_assign_params *params, const uint8_t *job_qv01, slr_umi_assign_rec *rec);

/* ---- S4: Illumina-guided barcode / UMI search ------------------------------------------------- */

/* Replaces, for a batch of reads, the offset loop of IlluminaUMIanalyzer.findUMI (F!com/rw/umifinder/analyzers/
 * IlluminaUMIanalyzer.class, IlluminaUMIanalyzer.java:L89-L136) or of IlluminaBarcodeAnalyzer.testBarcodes (F!…/
 * IlluminaBarcodeAnalyzer.class, IlluminaBarcodeAnalyzer.java:L272-L304) — one BCUMIEDtesterBase.matchesSeqEditDistance run
 * per offset (F!com/rw/nuc/encoding/TwoBit/ed/BCUMIEDtesterBase.class, BCUMIEDtesterBase.java:L82-L203) with the
 * checkMatchWithTestSets of UMInucTwoBitPerBaseEDtester (…java:L52-L67) or BCnucTwoBitPerBaseEDtester (…java:L72-L92) — and
 * the sorted().distinct() reduction of the collected list (IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI,
 * …java:L52-L60; testBarcodes L336-L339).  The host keeps the Needleman alignment of the two survivors and all flags. */
typedef struct slr_guided_sets slr_guided_sets;    /* device-resident candidate sets */

#define SLR_G_W_GENE  1u   /* entry (or an ancestor node) was found in the per-gene list: BARCODE_FOUND_FOR_GENE_OR_GENOMIC_REGION */
#define SLR_G_W_ALL   2u   /* BC_ONLY_FOUND_IN_ALL_PASSED_10xBCs */
#define SLR_G_W_EMPTY 4u   /* BC_IN_EMPTY_DROPS */
#define SLR_G_EXCEPTION 1u /* the Java would have thrown for this read (N in a window, non-IUPAC char, slice too short) */
#define SLR_G_TABLE_FULL 2u /* the device's visited table overflowed: the record is invalid (sized 4x above the measured maximum; never observed) */

typedef struct {
    uint64_t seq[2];       /* sequences of the first two entries of the sorted, distinct match list (2-bit packed) */
    int8_t   n_sub[2], n_ins[2], n_del[2], offset[2];   /* nSubstitutions / nInsertions / nDeletions / startOffsetFromPredicted */
    uint8_t  where[2];     /* SLR_G_W_* bits of findingErrorFlag (0 in the UMI flavour) */
    uint8_t  n_distinct;   /* min(size of the distinct list, 2): 0 = UMI_NOT_FOUND / no hit, 2 = a second-best match exists */
    uint8_t  flags;        /* SLR_G_EXCEPTION */
    int32_t  n_raw;        /* size of the raw list (matchLList.size()) */
    int32_t  min_err_gene; /* min getNErrors() over entries with the GENE bit (testBarcodes L312-L313), INT32_MAX = none */
    int32_t  pad;
} slr_guided_result;       /* 40 bytes */

typedef struct {
    uint64_t seq;
    int8_t   n_sub, n_ins, n_del, offset;
    uint8_t  where;        /* SLR_G_W_* */
    uint8_t  level;        /* currentlevel of the probed node */
    uint16_t pad;
} slr_guided_hit;          /* 16 bytes: one entry of the raw list, in list order */

/* Candidate sets.  group_keys / group_offsets: CSR of the groups (UMI flavour: the UMIs of one (gene, cell) =
 * IlluminaOneGeneOneCellData; BC flavour: the cell barcodes of one gene or genomic region = BarcodesMap), 2-bit packed.
 * BC flavour only: all_keys = All10xselectedCells searched while currentlevel <= all_ed (maxEDtoCheckBCAll10xBCs), NULL when
 * checkAllassignedBarcodes is false; empty_keys = EmptyDropBarcodes searched while currentlevel <= empty_ed
 * (maxEDtoCheckBCEmptyDrops), NULL when checkEmptyDrops is false.  seq_len = umi_length or cell_bc_length (<= 16). */
int  slr_guided_sets_create(slr_ctx *ctx, const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups,
                            const uint64_t *all_keys, int64_t n_all, int all_ed, const uint64_t *empty_keys, int64_t n_empty,
                            int empty_ed, int bc_flavour, int seq_len, slr_guided_sets **out);
void slr_guided_sets_destroy(slr_guided_sets *s);

/*   plusminus  umi_posplusminus / bc_posplusminus: offsets 0,-1,+1,… in that order
 *   post_len   bases handed to the tester as postUMIseq / postBCseq (UMI: ed + posplusminus + 2, the caller pads a short read
 *              with 'A' like findUMI L118-L124; BC: 10), ed + 1 <= post_len, seq_len + post_len <= 32
 *   bailout    umi_bailout_afterED / cell_BC_bailout_after_ED, < 0 = null
 *   slices     n * stride bytes, ASCII, STRANDED orientation (getSeqRevComp for 3' reads); window of offset i =
 *              slice[anchor+i, +seq_len), its post sequence the post_len bases that follow; slice_len <= 32
 *   group_id   candidate group of read i (out of range = no group: BC flavour searches only the global lists)
 *   ed         per read: maxEDdyn (slr_dyn_max_ed) or the fixed edit distance, 0..4
 *   raw_out    optional (NULL): the first raw_cap entries of every read's raw list, n * raw_cap records */
int  slr_guided_match(slr_ctx *ctx, const slr_guided_sets *s, int plusminus, int post_len, int bailout, const uint8_t *slices,
                      int stride, int slice_len, const int32_t *anchor, const int32_t *group_id, const int32_t *ed, int64_t n,
                      slr_guided_result *out, slr_guided_hit *raw_out, int raw_cap);
int  slr_guided_match_dev(slr_ctx *ctx, const slr_guided_sets *s, int plusminus, int post_len, int bailout,
                          const uint8_t *d_slices, int stride, int slice_len, const int32_t *d_anchor, const int32_t *d_group_id,
                          const int32_t *d_ed, int max_ed, int64_t n, slr_guided_result *d_out, slr_guided_hit *d_raw_out,
                          int raw_cap, void *stream);

/* DynamicEditDistances.getmaxED (F!com/rw/parameters/DynamicEditDistances.class, DynamicEditDistances.java:L93-L98): the largest
 * edit distance e whose max_candidates[e] >= count * (2 * plusminus + 1), capped at `cap` (< 0 = null).  max_candidates = one
 * <errorpercent> column of bcMaxEditDistances.xml / umiMaxEditDistances.xml.  Returns -1 when no entry qualifies (the Java
 * throws NoSuchElementException).  Pure host arithmetic. */
int  slr_dyn_max_ed(const int64_t *max_candidates, int n_ed, int count, int plusminus, int cap);

/* ---- several GPUs behind one caller (slr_init(n_devices, device_ids), slr_counts_reduce; §8e) ----------------------- */

/* The reference is ONE JVM (WorkerReadscanner.java:L186-L204: both worker pools live in the process that owns the FASTQ reader and the
 * writers), so the 8 GPUs of a box must be reachable from a single caller.  slr_multi owns one context per device; every slr_multi_* call
 * cuts its batch into contiguous shares, drives each device from its own host thread through the single-device entry point above and
 * returns when all shares are done.  Results are positional, exactly as from one device.
 *   n_devices <= 0 = every visible device; device_ids NULL = 0 .. n_devices-1; n_streams as slr_ctx_create (per device) */
typedef struct slr_multi slr_multi;
typedef struct slr_multi_table slr_multi_table;
int  slr_multi_create(int n_devices, const int *device_ids, int n_streams, slr_multi **out);
void slr_multi_destroy(slr_multi *m);
int  slr_multi_n_devices(const slr_multi *m);
slr_ctx *slr_multi_ctx(slr_multi *m, int i);                  /* the context of device i (borrowed), for the *_dev / session entries */
int  slr_multi_peer_access(const slr_multi *m);               /* 1 = every device reads every other device's HBM (NVLink / NVSwitch) */
/* BarcodesMapForBCfinding replicated on every device (the list is small: 3 M barcodes = 160 MB of tables) */
int  slr_multi_bc_table_create(slr_multi *m, const uint64_t *barcodes2bit, const int32_t *rank, int64_t n, int bc_len, slr_multi_table **out);
void slr_multi_bc_table_destroy(slr_multi_table *t);
slr_bc_table *slr_multi_bc_table_replica(slr_multi_table *t, int i);
/* Parser.assignBarcode / the pass-1 exact lookup for a batch spread over all devices (arguments as slr_bc_assign / slr_bc_exact) */
int  slr_multi_bc_assign(slr_multi *m, const slr_multi_table *t, int ed_max, int plusminus, int three_prime, const uint8_t *slices, int stride,
                         int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out);
int  slr_multi_bc_exact(slr_multi *m, const slr_multi_table *t, int three_prime, const uint8_t *slices, int stride, int slice_len,
                        const int32_t *lens, const int32_t *anchor, int64_t n, slr_bc_result *out);
/* assignedBarcodes2ndPass over all devices: the replicas' counters are summed on device 0 by one kernel that loads the peers' arrays over
 * NVLink (staged copies when peer access is unavailable), then copied out — the run's single cross-device reduction */
int  slr_multi_bc_counts_read(slr_multi *m, slr_multi_table *t, int64_t *counts_out);
int  slr_multi_bc_counts_reset(slr_multi *m, slr_multi_table *t);
/* UMI seams: the (cell, region) jobs are dealt to the devices in contiguous runs of WHOLE jobs balanced by n^2, so no job is ever cut and no
 * cross-device merge exists (arguments as slr_umi_dist / slr_umi_cluster / slr_umi_assign; records and matrices positional) */
int  slr_multi_umi_dist(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int32_t *out,
                        const int64_t *out_offsets);
int  slr_multi_umi_cluster(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs, int ed,
                           const uint8_t *member, const int32_t *rank, slr_umi_cluster_rec *rec);
int  slr_multi_umi_assign(slr_multi *m, const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                          const slr_umi_assign_params *params, const uint8_t *job_qv01, slr_umi_assign_rec *rec);

/* ---- misc ------------------------------------------------------------------------------------------- */
const char *slr_last_error(void);
int  slr_abi_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t slr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
