/*
 * sicelore_host.h — the HOST-SIDE entry points of libsicelore_gpu.so: the callers' side of three seams, for callers without a JVM.
 *
 * Everything here is plain host arithmetic (no device, no slr_ctx): the reference does this work in Java on one thread between or around
 * the loops that sicelore_gpu.h moves to the GPU, and a JVM caller simply keeps its own classes.  A C / C++ / Python driver gets the same
 * results from these functions, each pinned against the reference's own class files (tests/golden/ref_grouper.npz, ref_jobs.npz,
 * ref_needleman.npz, ref_usedlist.npz).  Conventions as in sicelore_gpu.h (caller-owned buffers, 0 or a negative SLR_E_* code,
 * slr_last_error()).  Kept in its own header so that the kernels' translation units — which include sicelore_gpu.h for the record
 * layouts — do not depend on it.
 */
#ifndef SICELORE_HOST_H
#define SICELORE_HOST_H
#include <stdint.h>
#include "sicelore_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SLR_E_REFERENCE_THROWS (-6)   /* the reference itself throws on this input (the message names the exception and the source line) */

/* ---- between the passes: from the pass-1 counts and the collision records to the used-barcode list of pass 2 (host arithmetic, no device) ---- */

/* UsedBarcodesListData.filterLowCounts as finalizeData calls it (F!…/UsedCellBCListGenerator$UsedBarcodesListData.class,
 * UsedCellBCListGenerator.java:L359-L363, L391-L392): keep_out[i] = counts[i] > 2.0f * record_count / 5000000.0f (float) && counts[i] > 1.
 * counts = unfilteredUsedBarcodeMap (slr_bc_counts_read after slr_bc_exact), record_count = reads scanned in pass 1.  The kept barcodes are the
 * list the collision tester runs on (slr_bc_table_create + slr_bc_collide of the list against itself). */
int  slr_bc_used_filter_low_counts(const int32_t *counts, int64_t n, int64_t record_count, uint8_t *keep_out);

#define SLR_UL_ORDER_UNPIN 1u  /* a java.util.HashMap bin reached 9 entries at >= 64 bins (a JDK tree bin): the iteration order that decides chains of removals is not reproduced */
#define SLR_UL_RANK_TIES   2u  /* kept barcodes with equal counts: their relative ranks follow fastutil's table order in the reference, input order here */
/* Replaces BarcodeDatasetColissionTester.generateColissionMergedBCmap (F!…/BarcodeDatasetColissionTester.class, …java:L158-L203) and the rank
 * assignment of WorkerReadscanner.java:L264-L270.  barcodes / counts: the count-filtered list; collide[i]: the record slr_bc_collide returned
 * for barcodes[i] against this same list; min_count_fold = minCountFold (config.xml:61), merge_ed = mergeBCsED (null = --bcEditDistance),
 * cells_fold = cellsWithReadsnFoldBelowMaxToKeep (config.xml:27).  A barcode B removes every collider c (ED <= merge_ed) with
 * counts[c] < counts[B] / min_count_fold — visited in the JDK HashMap's iteration order, and a barcode that has itself been removed removes
 * nobody; the survivors with counts >= max / cells_fold are kept.  keep_out[n]; rank_out[n] (may be NULL): 1 = most reads, 0 = dropped;
 * *flags_out (may be NULL): SLR_UL_*. */
int  slr_bc_used_merge_collisions(const uint64_t *barcodes, const int32_t *counts, const slr_collide_result *collide, int64_t n,
                                  int min_count_fold, int merge_ed, int cells_fold, uint8_t *keep_out, int32_t *rank_out, uint32_t *flags_out);

/* ---- seam S4, after slr_guided_match: MORE_THAN_ONE_MATCH ----------------------------------------------- */

/* The alignment comparison that decides MORE_THAN_ONE_MATCH (host arithmetic on the two survivors of a record, no device): replaces the two
 * NeedlemanWunsch alignments of IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI (…java:L66-L79; T!com/rw/nuc/alignment/needleman/
 * NeedlemanWunsch.class) and NeedlemanMatch.countNeedlemanErrorsInRead (F!com/rw/nanopore/analyzers/NeedlemanMatch.class, …java:L68-L86).
 * scores = the reference's NeedlemanScores (NeedlemanParameters.umi / .bc), NULL = its defaults (-4, -5, -5, -5, -5, -5, 5). */
typedef struct { int32_t leading_gap_1, leading_gap_2, trailing_gap_1, trailing_gap_2, indel, mismatch, match; } slr_needleman_scores;
#define SLR_G_NO_SECOND INT32_MIN      /* slr_guided_mismatch_diff: the record has no second-best entry (or is flagged) */
/* one alignment: candidate (template) vs read window, both 2-bit packed, len <= 32; counts_out[4] = insertionsNeedleman, deletionsNeedleman
 * (gaps at the end of the read row are not counted), substitutionsNeedleman, their sum (getNerrorsNeedleman) */
int  slr_needleman_errors(uint64_t template2bit, uint64_t read2bit, int len, const slr_needleman_scores *scores, int32_t *counts_out);
/* per record of slr_guided_match (same slices / anchor / seq_len): diff_out[i] = nMismatchDiffBestvsSecondBest = errors(second) - errors(best),
 * each entry aligned to the window at anchor + offset it was found from (its unMutatedSeq); SLR_G_NO_SECOND when n_distinct < 2.
 * diff == 0 <=> the reference sets MORE_THAN_ONE_MATCH (…java:L80-L86), i.e. the read counts as not found (IlluminaUMIanalyzer.java:L203-L220). */
int  slr_guided_mismatch_diff(const slr_guided_result *res, int64_t n, const uint8_t *slices, int stride, int slice_len, const int32_t *anchor,
                              int seq_len, const slr_needleman_scores *scores, int32_t *diff_out);

/* ---- host side of the clustering seam: forming the (cell, region) jobs ------------------------------- */
/* Plain host code (no device, no context): the reference does this on its BAM reader thread, a JVM caller keeps its own classes; these entry
 * points exist for callers WITHOUT a JVM, so that they feed slr_umi_assign with the same jobs.  Pinned against the reference's class files
 * (tests/golden/ref_grouper.npz, ref_jobs.npz). */
typedef struct slr_grouper slr_grouper;   /* the process-wide state of ReadGrouper: MAX_GENOME_DISTANCE_FOR_SAME_GENOMIC_REGION + the static region counter */

/* ReadGrouper.setMaxGenomeDistance (config.xml:247 max_GenomeDistance_forGrouping, default 500) + ReadGrouper$Cluster.CURRENT_GENOMIC_REGION_ID
 * (ReadGrouper.java:L460; 0 at JVM start).  Like the reference's static state it is meant for ONE reader thread: calls on the same grouper must
 * not overlap (different groupers are independent). */
int  slr_grouper_create(int max_genome_distance, int64_t first_region_id, slr_grouper **out);
void slr_grouper_destroy(slr_grouper *g);
int64_t slr_grouper_next_region_id(const slr_grouper *g);

/* Replaces ReadGrouper.groupSams (F!com/rw/umifinder/bamreaders/ReadGrouper.class, ReadGrouper.java:L82-L230; caller BamReader.run,
 * BamReader.java:L134-L145) for one chunk of n SAM records in BAM order.  position[i] = ReadScanData.positionOnGenomeForClustering (read only
 * where has_position[i] != 0; has_position NULL = every read has one), flags[i] = SAMRecord.getFlags() (bit 16 = reverse strand),
 * region_io[i] = NanoporeRead.genomicRegionNmber, -1 = absent: the reads of every surviving cluster receive its number, the others keep
 * what they had (a read carried over from the previous chunk keeps that round's number).  *last_index_out: records [0, last_index] are the
 * grouped chunk the clustering stage receives; with keep_data_end the records behind it open the caller's next chunk (BamReader.java:L134), without
 * they are dropped (the reference returns an empty chunk).  n == 0: *last_index_out = -1, nothing is handed on (L82-L83).
 * SLR_E_REFERENCE_THROWS: the reference's NullPointerException at L173 (keep_data_end, one surviving cluster whose centre cache is empty). */
int  slr_grouper_group_sams(slr_grouper *g, const int32_t *position, const uint8_t *has_position, const int32_t *flags, int64_t n,
                            int keep_data_end, int64_t *region_io, int64_t *last_index_out);

/* Replaces UmiClustering.cluster up to the hand-over to its Submitter (F!…/clustering/UmiClustering.class, UmiClustering.java:L97-L118
 * groupDataByCellAndRegion, L135 size filter, L136-L142 split of oversized groups): the reads with valid[i] != 0 (a cell barcode AND a region
 * number; NULL = all) grouped by (cell_bc, region); groups of fewer than min_size reads (the reference: 2) are dropped; ram_reserved != 0 cuts a
 * group of n reads into ceil((float) n / sqrt(ram_reserved / 300)) consecutive parts of n / nChunks + 1 reads like the reference's memory bound
 * does (RAM_RESERVED, UmiClustering.java:L59; 0 = never split — the GPU has no such bound, but the split changes the clusters).
 * order_out (capacity n): read indices, job j = order_out[job_offsets_out[j] .. job_offsets_out[j + 1]) in input order; job_offsets_out has
 * capacity n + 1; jobs in ascending (cell_bc, region) order (the reference's map iteration order reaches no per-read result). */
int  slr_group_jobs(const uint64_t *cell_bc, const int64_t *region, const uint8_t *valid, int64_t n, int min_size, int64_t ram_reserved,
                    int64_t *order_out, int64_t *job_offsets_out, int64_t *n_jobs_out);

#ifdef __cplusplus
}
#endif
#endif
