/*
 * slr_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the SiCeLoRe 2.1 barcode / UMI edit-distance hot path, written from the
 * shipped bytecode (no Java source exists in the reference tree).  Citations use the survey's
 * notation:  F! = Jar/NanoporeBC_UMI_finder-2.1.jar,  T! = Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar,
 * (Foo.java:Lnnn) = original source line recovered from the LineNumberTable (tools/jdis.py).
 *
 * PARITY STATUS: the reference ships no tests / golden vectors for this path and there is no JVM in the build container, but its class
 * files are: oracle/minijvm.py (a JVM-subset interpreter) executes them unmodified and oracle/make_ref_vectors.py freezes the outputs in
 * tests/golden/ref_*.npz.  Pinned that way (tests/test_ref_vectors.py): the 2-bit primitives, limitedCompare, best-of-9 + packing,
 * calcEditDistances (UMI window slicing included), BarcodeMatchTester.doJob, Parser.assignBarcode, the pass-1 exact lookup, the
 * Illumina-guided testers, IlluminaUMIanalyzer.findUMI and IlluminaBarcodeAnalyzer.testBarcodes (one gene) with getBestAndSecondBCorUMI,
 * getmaxED.  JDK / third-party containers are shims there (java.util.HashSet iteration order = the JDK HashMap algorithm as modelled,
 * not executed).  Also: the two read-name examples of
 * /root/reference/README.md:400,452, an independent second restatement (oracle/pyref.py) and brute-force property checks (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link
 * or call this file.  The product (libsicelore_gpu.so) never does.
 */
#ifndef SLR_ORACLE_H
#define SLR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ED 8

/* ---------- 2-bit primitives (T!com/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase) ---------- */
uint64_t orc_pack2bit(const uint8_t *chars, int len, int *bad_char);          /* java:L183-L187 */
uint64_t orc_revcomp2bit(uint64_t seq, int len);                              /* java:L477-L484 */
void     orc_replace_deg(uint64_t seq, uint64_t out[4], int pos, int len);    /* java:L228-L234 */
void     orc_insert_deg(uint64_t seq, uint64_t out[4], int pos, int len);     /* java:L300-L310 */
uint64_t orc_delete_byte(uint64_t seq, int code4, int pos, int len);          /* java:L321-L327 */
int      orc_encode4bit(uint8_t c);   /* NucleicAcidByteCodeBase.ENCODE_MATRIX, 0xFF = unknown (java:L45-L78) */
int      orc_revcomp4bit(int code);   /* ONEBYTE_REVERSECOMP_MATRIX (java:L100-L133) */

/* ---------- barcode search set (stands in for fastutil LongSet.contains: membership only) ------ */
typedef struct orc_set orc_set;
orc_set *orc_set_new(const uint64_t *keys, int64_t n);
void     orc_set_free(orc_set *s);
int64_t  orc_set_find(const orc_set *s, uint64_t key);   /* index into keys[] or -1 */
int64_t  orc_set_size(const orc_set *s);

/* ---------- BarcodeMatchTester.doJob (F!...BarcodeMatchTester.java:L198-L244) ------------------ */
typedef struct {
    uint64_t read_seq;      /* OneMatch.readSeq = unmutated window */
    uint64_t bc;            /* OneMatch.matchingBC */
    int32_t  ed;            /* OneMatch.editDistance = DFS level of the hit */
    int32_t  offset;        /* OneMatch.offsetFromPredicted */
    int32_t  n_sub, n_ins, n_del;   /* counters exactly as the Java names them (INS op bumps nDel) */
} orc_match;

/* Returns number of matches (0..ed+1) in DISCOVERY order (= HashSet chain order, all share one hash),
 * or -1 if the Java would have thrown (invalid post code reached).  post4 = 4-bit codes, post_len<0 => null. */
int orc_match_tester(const orc_set *set, uint64_t seq, int len, int ed, int skip_full_matches,
                     int allow_indels, const uint8_t *post4, int post_len, int do_next_level_if_match,
                     int offset, orc_match *out, int64_t *n_probes);

/* ---------- Parser.assignBarcode (F!...Parser.java:L195-L315) ---------------------------------- */
#define ORC_F_ASSIGNED   1u   /* BC_FOUND */
#define ORC_F_EXCEPTION  2u   /* the Java would have thrown (index out of range / unknown char) */
#define ORC_F_TIE_UNPIN  4u   /* HashMap bin got treeified: iteration order not emulated */

typedef struct {
    uint64_t bc;            /* best.matchingBC if assigned else 0 */
    int32_t  ed;            /* best ED; -1 = no match at any offset */
    int32_t  ed_second;     /* second-best ED (distinct barcode), INT32_MAX = none */
    int8_t   offset;        /* best.offsetFromPredicted (assigned only) */
    int8_t   n_ins;         /* OneMatch.insertions */
    int8_t   n_del;         /* OneMatch.deletions */
    int8_t   n_sub;
    int32_t  rank;          /* CountsRank.rank of bc, -1 if unassigned */
    uint32_t flags;
} orc_bc_result;            /* 32 bytes, same layout as slr_bc_result */

void orc_assign_barcode(const orc_set *set, const int32_t *rank, int ed_max, int plusminus, int three_prime,
                        int bc_len, const uint8_t *slice, int slice_len, int anchor,
                        orc_bc_result *out, int64_t *n_probes);

/* batch, OpenMP over reads (used as the CPU baseline).  slices: n * stride bytes */
void orc_assign_barcode_batch(const orc_set *set, const int32_t *rank, int ed_max, int plusminus, int three_prime,
                              int bc_len, const uint8_t *slices, int stride, int slice_len, const int32_t *anchor,
                              int64_t n, orc_bc_result *out, int64_t *n_probes_total, int n_threads);

/* ---------- pass-1 exact lookup (F!...UsedCellBCListGenerator$Worker, UsedCellBCListGenerator.java:L206-L232) ---------- */
/* window at the predicted position only, no post sequence; found => flags = ASSIGNED, bc, ed = 0, rank */
void orc_exact_lookup_batch(const orc_set *set, const int32_t *rank, int three_prime, int bc_len, const uint8_t *slices, int stride,
                            int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, orc_bc_result *out);

/* ---------- pass-1 collision tester (F!...BarcodeDatasetColissionTester.java:L212-L229) ---------- */
typedef struct {
    uint64_t bc[2];         /* matchingBC of the ED-1 / ED-2 OneMatch */
    uint8_t  valid;         /* bit 0: ED-1 entry, bit 1: ED-2 entry */
    uint8_t  n_sub[2], n_ins[2], n_del[2];
    uint8_t  pad;
} orc_collide_result;       /* 24 bytes, same layout as slr_collide_result */

/* one BarcodeMatchTester(seq, ed, skipFullMatches=true, allowIndels=true, set, offset 0, len, postSeq=null,
 * doNextLevelIfMatchFound=false).call() per query (L215-L222); OpenMP over queries */
void orc_collide_batch(const orc_set *set, int ed, int bc_len, const uint64_t *queries, int64_t n, orc_collide_result *out,
                       int64_t *n_probes_total, int n_threads);

/* ---------- Illumina-guided search engine ----------------------------------------------------------
 * F!com/rw/nuc/encoding/TwoBit/ed/BCUMIEDtesterBase.matchesSeqEditDistance (BCUMIEDtesterBase.java:L82-L124, L136-L203),
 * checkMatchWithTestSets of UMInucTwoBitPerBaseEDtester (java:L52-L67) and BCnucTwoBitPerBaseEDtester (java:L72-L92),
 * goNextEDlevel with bailoutIfFoundAfterED (NucTwoBitPerBaseEDtesterBase.java:L133-L144), the offset loops of
 * IlluminaUMIanalyzer.findUMI (java:L89-L136) / IlluminaBarcodeAnalyzer.testBarcodes (java:L272-L304) and the
 * sorted().distinct() reduction of IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI (java:L52-L60, comparators L106-L110, L143-L144). */
#define ORC_W_GENE  1u    /* BarcodeFindingFlag.BARCODE_FOUND_FOR_GENE_OR_GENOMIC_REGION (own or inherited from an ancestor node) */
#define ORC_W_ALL   2u    /* BC_ONLY_FOUND_IN_ALL_PASSED_10xBCs */
#define ORC_W_EMPTY 4u    /* BC_IN_EMPTY_DROPS */

typedef struct {
    const orc_set *group;  /* UMI flavour: umis of the (gene, cell); BC flavour: cellBcsForGene (NULL when empty, java:L299) */
    const orc_set *all;    /* allPassed10xBCs, NULL = checkAllassignedBarcodes false (UMI flavour: NULL) */
    int all_ed;            /* maxEDtoCheckBCAll10xBCs */
    const orc_set *empty;  /* outOfCellsBarcodes, NULL = checkEmptyDrops false or no empty-drop list */
    int empty_ed;          /* maxEDtoCheckBCEmptyDrops */
    int bc_flavour;        /* 1 = BCnucTwoBitPerBaseEDtester (a group hit puts the GENE bit on the node), 0 = UMInucTwoBitPerBaseEDtester */
} orc_guided_sets;

typedef struct {
    uint64_t seq;                          /* the matching (mutated) sequence */
    int8_t   n_sub, n_ins, n_del, offset;  /* counters as the Java names them; startOffsetFromPredicted */
    uint8_t  where;                        /* ORC_W_* bits of findingErrorFlag */
    uint8_t  level;                        /* currentlevel of the probed node (root: 1 with 0 errors) */
    uint16_t pad;
} orc_guided_hit;                          /* 16 bytes */

/* one tester (one window): the matchingList in list order.  Returns its size (the first `cap` entries are stored),
 * -1 if the Java would have thrown.  bailout < 0 = null.  post4: 4-bit codes (non-NULL in both callers). */
int64_t orc_guided_tester(const orc_guided_sets *sets, uint64_t seq, int len, int ed, int allow_indels, const uint8_t *post4,
                          int post_len, int bailout, int offset, orc_guided_hit *out, int64_t cap, int64_t *n_probes);

#define ORC_G_EXCEPTION 1u
typedef struct {
    uint64_t seq[2];                       /* first two entries of the sorted, distinct list */
    int8_t   n_sub[2], n_ins[2], n_del[2], offset[2];
    uint8_t  where[2];
    uint8_t  n_distinct;                   /* min(size of the distinct list, 2) */
    uint8_t  flags;                        /* ORC_G_EXCEPTION */
    int32_t  n_raw;                        /* size of the raw list over all offsets */
    int32_t  min_err_gene;                 /* min getNErrors() over entries carrying the GENE bit (testBarcodes L312-L313), INT32_MAX none */
    int32_t  pad;
} orc_guided_result;                       /* 40 bytes, same layout as slr_guided_result */

/* one query = one read against one candidate group: offsets 0,-1,+1,... (sorted by |i|, stable), window =
 * slice[anchor+i, +len), post = slice[anchor+i+len, +post_len) (stranded orientation, ASCII), then the reduction.
 * bc_flavour selects the comparator with scoreWhereFound.  raw_out (optional): first raw_cap entries of the raw list. */
void orc_guided_query(const orc_guided_sets *sets, int bc_flavour, int len, int ed, int plusminus, int bailout, int post_len,
                      const uint8_t *slice, int slice_len, int anchor, orc_guided_result *out, orc_guided_hit *raw_out,
                      int64_t raw_cap, int64_t *n_probes);

/* batch over queries (OpenMP).  group_keys/group_offsets: CSR of the candidate groups (2-bit packed), query i searches group
 * group_id[i] at edit distance ed[i]; all_keys / empty_keys: the two global lists of the BC flavour (NULL = unused). */
void orc_guided_batch(const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups, const uint64_t *all_keys,
                      int64_t n_all, int all_ed, const uint64_t *empty_keys, int64_t n_empty, int empty_ed, int bc_flavour, int len,
                      int plusminus, int bailout, int post_len, const uint8_t *slices, int stride, int slice_len,
                      const int32_t *anchor, const int32_t *group_id, const int32_t *ed, int64_t n, orc_guided_result *out,
                      orc_guided_hit *raw_out, int64_t raw_cap, int64_t *n_probes_total, int n_threads);

/* ---------- UMI distances (F!com/rw/clustering/ClusteringEditDistanceBase.java:L297-L350) ------ */
int     orc_limited_compare(const uint8_t *left, int n, const uint8_t *right, int m, int threshold); /* apachemod/LevenshteinDistance.java:L220-L283 */
int32_t orc_umi_best9(const uint8_t *a, const uint8_t *b, int umi_len);    /* a,b: umi_len+2 4-bit codes (window -1..+1) */
int32_t orc_umi_transpose(int32_t packed);                                  /* BestEditDistance.getTransposedCopy L458 */
int32_t orc_umi_equality(void);                                             /* EQUALITYMATRIX result (L90-L92) */
/* generateDistanceMatrix for one (cell, region) job: n reads, out = n*n row-major packed ints */
void    orc_umi_matrix(const uint8_t *umis, int stride, int umi_len, int64_t n, int32_t *out);
void    orc_umi_matrix_batch(const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                             int32_t *out, const int64_t *out_offsets, int n_threads);


/* ---------- neighbour-set clustering (F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.java:L175-L219) ------
 * clusterLocal on one job's n x n packed matrix.  member: NULL = all reads are in `indices`.  order: NULL, or the
 * keys of possibleClusters in the map's iteration order (n_order entries, job-local indices; non-keys are skipped).
 * rec[c] = { |N(c)|, chosen entry or -1, |N(entry)|, number of entries tied for the maximum }. */
typedef struct { int32_t n_neighbours, best_key, best_count, n_ties; } orc_cluster_rec;
void    orc_umi_cluster(const int32_t *matrix, int64_t n, int ed, const uint8_t *member, const int32_t *order, int64_t n_order,
                        orc_cluster_rec *rec);
/* all jobs; rank (per read, job-local iteration rank of the key, NULL = ascending index) is turned into `order` per job */
void    orc_umi_cluster_batch(const int32_t *matrices, const int64_t *job_offsets, const int64_t *out_offsets, int64_t n_jobs, int ed,
                              const uint8_t *member, const int32_t *rank, orc_cluster_rec *rec, int n_threads);

/* ---------- clustering + UMI assignment of a job (oracle/slr_oracle_assign.c) ------------------------------------------------
 * ClusterOneHierarchical.call (F!com/rw/umifinder/analyzers/clustering/ClusterOneHierarchical.java:L61-L217): the clusterer of every job of at
 * most 100 reads (UmiClustering.java:L240) — LingPipe complete link (CompleteLinkClusterer.java:L146-L237) / single link on the reads that
 * have a neighbour, Dendrogram.partitionDistance, the depth rule, OneUmiCluster.setClusterCenter (OneUmiCluster.java:L49-L65) and the
 * per-read values ClusterOneBase.setSamflagsAndStatsForClustered writes (ClusterOneBase.java:L118-L168). */
#define ORC_UA_ASSIGNED  1u   /* the read is in a cluster of the final list: setSamflagsAndStatsForClustered runs for it */
#define ORC_UA_SKIPPED   2u   /* its cluster failed the depth rule: flagDontUMIassignRecords (UMI_CLUSTERING_SKIPPED_HIGHCOMPLEXITY) */
#define ORC_UA_TIE_UNPIN 4u   /* the job's result can depend on the JVM's identity-hash order (ObjectToSet's HashSet<PairScore>) */
#define ORC_UA_DEEP      8u   /* job of more than max_hier reads: ClusterOne_MyClustering's (clustered by orc_umi_assign_myclust when params.deep) */
typedef struct {
    int32_t ed_complete;      /* umi_completelinkclusteringED (config.xml:270) */
    int32_t ed_single;        /* umi_singlelinkclusteringED (config.xml:272) */
    int32_t single_threshold; /* complexity_threshold_for_switch_to_single_link_clustering (config.xml:278) */
    int32_t fold_depth;       /* foldDepthBelowMaxDiscardForClustering (UMIparameters.java:L118: 50) */
    int32_t max_hier;         /* largest job ClusterOneHierarchical gets (UmiClustering.java:L240: 100) */
    int32_t deep;             /* 1: larger jobs run ClusterOne_MyClustering.call (records carry ORC_UA_DEEP too); 0: they are only flagged ORC_UA_DEEP */
} orc_assign_params;
typedef struct {
    int32_t  center;          /* job-local index of OneUmiCluster.getCenter() of the read's cluster; -1 = not clustered */
    int8_t   u1;              /* UMI_ED: distanceNonReducedSet(center, read) */
    int8_t   u2;              /* UMI_ED_SECOND_BEST_MATCH: least distance to a read outside the cluster; -1 = tag not written */
    int8_t   pos2;            /* matrix[center][read].getPos2(): 0 MINUSONE, 1 ZERO, 2 PLUSONE */
    int8_t   off_mean;        /* offsetcentermean of the cluster (ClusterOneHierarchical.java:L143-L147) */
    uint16_t flags;           /* ORC_UA_* */
    uint16_t cluster_size;
    int32_t  n_clusters;      /* cluster_list.size() of the job */
} orc_assign_rec;             /* 16 bytes, same layout as slr_umi_assign_rec */
void orc_umi_assign_hier(const int32_t *matrix, int64_t n, const orc_assign_params *P, int qv01, orc_assign_rec *rec);
int  orc_fu_set_ops(const int *keys, int k, const int *victims, int nv, int max_key, int *order_before, int *order_after);   /* test hook */
void orc_umi_assign_myclust(const int32_t *matrix, int64_t n, const orc_assign_params *P, int qv01, orc_assign_rec *rec);
void orc_umi_assign_batch(const int32_t *matrices, const int64_t *job_offsets, const int64_t *out_offsets, int64_t n_jobs,
                          const orc_assign_params *P, const uint8_t *job_qv01, orc_assign_rec *rec, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
