// slr_kernels.h — launchers of the CUDA kernels (internal to libsicelore_gpu.so)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "slr_table.cuh"
#include "../../include/sicelore_gpu.h"

// need_post = 0: exact lookup of the window only (pass-1 used-barcode counting), no post sequence is taken
cudaError_t slr_launch_bc_assign(const SlrTableDev &tab, int ed_max, int plusminus, int three_prime, int need_post, const uint8_t *d_slices,
                                 int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor, long long n,
                                 slr_bc_result *d_out, unsigned long long *d_work, cudaStream_t stream);   // d_work: 8 bytes, this launch's own

// d_scratch: slr_umi_scratch_bytes(n_reads) bytes of device memory (8-byte aligned) that stay untouched until the launch
// has finished on `stream`; three kernels are enqueued (SLR_UMI_LAUNCHES)
constexpr int SLR_UMI_LAUNCHES = 3;
size_t slr_umi_scratch_bytes(long long n_reads);
cudaError_t slr_launch_umi_dist(const uint8_t *d_umis, int stride, int umi_len, const long long *d_job_offsets, long long n_jobs,
                                long long n_reads, int32_t *d_out, const long long *d_out_offsets, void *d_scratch, cudaStream_t stream);

// neighbour counts + best cluster key on the matrices slr_launch_umi_dist wrote (umi_cluster.cu); d_counts: n_reads int32
constexpr int SLR_UMI_CLUSTER_LAUNCHES = 4;
// d_range: 16 bytes of device memory owned by this launch (the read range of the deep jobs, filled by the first kernel)
// d_rowjob: the job of every read as slr_launch_umi_dist left it in its scratch (slr_umi_scratch_rowjob), or NULL (binary search)
const int32_t *slr_umi_scratch_rowjob(const void *d_scratch, long long n_reads);
cudaError_t slr_launch_umi_cluster(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                   long long n_reads, int ed, const uint8_t *d_member, const int32_t *d_rank, const int32_t *d_rowjob,
                                   int32_t *d_counts, slr_umi_cluster_rec *d_out, unsigned long long *d_range, cudaStream_t stream);

// ClusterOneHierarchical.call for every job of at most P.max_hier reads + the per-read assignment values (umi_assign.cu); three kernels.
// d_scratch: slr_umi_assign_scratch(n_jobs, deep_words) bytes — job lists + the working arrays of the jobs above max_hier (deep_words =
// sum of slr_umi_assign_deep_words(n) over those jobs; 0 or P.deep == 0: they are only flagged SLR_UA_DEEP);
// d_rowjob as for slr_launch_umi_cluster (NULL: binary search)
constexpr int SLR_UMI_ASSIGN_LAUNCHES = 3;
constexpr int SLR_UMI_ASSIGN_DEEP_LAUNCHES = 3;
constexpr int SLR_UA_DEEP_SMALL = 1024;        // deep jobs up to this size run on one CTA, up to SLR_UA_DEEP_MEDIUM on a cluster of 8 CTAs,
constexpr int SLR_UA_DEEP_MEDIUM = 4096;       // larger ones on the whole GPU (cooperative grid), one job after the other
size_t slr_umi_assign_scratch(long long n_jobs, long long deep_words);
// 32-bit words of working arrays umi_assign_deep.cu needs for a job of n reads (kept in step with carve() there): ~70 n words of lists and
// hash tables + the n x n threshold bit matrix (n^2 / 32 words)
SLR_HD long long slr_umi_assign_deep_words(long long n) { return (64 + 52 * (n + 4) + 2 * (3 * n + 64) + 2 * (4 * n + 64) + n * ((n + 31) / 32) + 1) & ~1ll; }
cudaError_t slr_launch_umi_assign(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets, long long n_jobs,
                                  long long n_reads, const slr_umi_assign_params &P, const uint8_t *d_job_qv01, const int32_t *d_rowjob,
                                  slr_umi_assign_rec *d_rec, void *d_scratch, size_t scratch_bytes, cudaStream_t stream);
cudaError_t slr_launch_umi_assign_deep(const int32_t *d_mat, const long long *d_job_offsets, const long long *d_out_offsets,
                                       const slr_umi_assign_params &P, const uint8_t *d_job_qv01, slr_umi_assign_rec *d_rec,
                                       const int32_t *d_list_small, const long long *d_off_small, const unsigned int *d_count_small,
                                       const int32_t *d_list_big, const long long *d_off_big, const unsigned int *d_count_big, int *d_words,
                                       long long max_jobs, cudaStream_t stream);

// a range of the caller's CSR job offsets (d_raw: n_jobs + 1 entries) rebased on the device: d_joff = raw - r0, d_ooff = exclusive prefix
// sum of the squared job sizes (both n_jobs + 1 entries); d_tmp: slr_umi_rebase_tmp_bytes(n_jobs); three kernels
constexpr int SLR_UMI_REBASE_LAUNCHES = 3;
size_t slr_umi_rebase_tmp_bytes(long long n_jobs);
cudaError_t slr_launch_umi_rebase(const long long *d_raw, long long n_jobs, long long r0, long long *d_joff, long long *d_ooff, void *d_tmp,
                                  cudaStream_t stream);

cudaError_t slr_launch_bc_collide(const SlrTableDev &tab, int ed_max, const unsigned long long *d_queries, long long n,
                                  slr_collide_result *d_out, cudaStream_t stream);

// Illumina-guided search (guided_match.cu).  d_vis: slr_guided_vis_bytes(max_ed) bytes, zeroed by the launcher; every ed[i] <= max_ed
#include "guided_core.cuh"
size_t slr_guided_vis_bytes(int max_ed, int *warps_out);
cudaError_t slr_launch_guided_match(const SlrGuidedSetsDev &S, int L, int plusminus, int post_len, int bailout, const uint8_t *d_slices,
                                    int stride, int slice_len, const int32_t *d_anchor, const int32_t *d_group_id, const int32_t *d_ed,
                                    int max_ed, long long n, slr_guided_result *d_out, slr_guided_hit *d_raw, int raw_cap, void *d_vis,
                                    unsigned long long *d_work, cudaStream_t stream);
