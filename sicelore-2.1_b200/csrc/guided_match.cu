// guided_match.cu — Illumina-guided barcode / UMI search for sm_100a.
//
// One warp per read (query).  For every offset window the warp walks the reference's depth-first enumeration
// (BCUMIEDtesterBase.matchesSeqEditDistance, F!…/TwoBit/ed/BCUMIEDtesterBase.class, BCUMIEDtesterBase.java:L82-L124) in the
// reference's own order, because the match list is ordered and the visited set / the bailout depend on what ran before:
//   * the deque lives in shared memory (per warp, LIFO like ArrayDeque.add / pollLast);
//   * a node below the last level runs ONE position per pop: its <= 9 children are generated, tested against the visited
//     set, probed and pushed by 9 lanes (ballot + prefix ranks keep the creation order);
//   * a node of the last level (the bulk: ~(9L)^(ed-1) of them) runs ALL its positions as one batch of 9·L children, 32 per
//     warp step — its children are never pushed, and the only change of the visited set in between is the node's own
//     sequence after its first position (the "self rule" below);
//   * hits are folded, in list order, into the first two entries of the consumers' sorted().distinct() list.
// The visited set is a per-warp table in global memory (stamped slots, never cleared between windows); the candidate
// sets are small open-addressing tables that stay in L1/L2.  Integer work only: no tensor cores.
#include <cuda_runtime.h>
#include "guided_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int G_WARPS = 8;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct GWarpShared {
    SlrGNode stack[SLR_G_STACK];
    uint8_t codes[32];
    uint32_t peq[2][4][32];                    // match masks of the lane's two filter candidates (slr_g_peq), per-read constants
};

// list order bookkeeping of one hit (all lanes hold the same state); lane 0 writes the optional raw record
__device__ __forceinline__ void g_record(SlrGTop2 &T, int bc_flavour, uint32_t seq, uint32_t cmeta, int level, int offset, uint32_t where,
                                         slr_guided_hit *raw, int raw_cap, int lane)
{
    if (raw && T.n_raw < raw_cap && lane == 0) {
        slr_guided_hit h;
        h.seq = seq; h.n_sub = (int8_t)slr_g_nsub(cmeta); h.n_ins = (int8_t)slr_g_nins(cmeta); h.n_del = (int8_t)slr_g_ndel(cmeta);
        h.offset = (int8_t)offset; h.where = (uint8_t)(where & 7u); h.level = (uint8_t)level; h.pad = 0;
        raw[T.n_raw] = h;
    }
    slr_g_top2_add(T, bc_flavour, seq, cmeta, offset, where);
}

// VIS_SMEM: every ed of the batch is <= 2, the visited tables (512 slots per warp) live in shared memory
template <bool VIS_SMEM>
__global__ void __launch_bounds__(G_WARPS * 32, VIS_SMEM ? 4 : 3)
guided_match_kernel(SlrGuidedSetsDev S, int L, int plusminus, int post_len, int bailout, const uint8_t *__restrict__ slices, int stride,
                    int slice_len, const int32_t *__restrict__ anchor, const int32_t *__restrict__ group_id,
                    const int32_t *__restrict__ ed_arr, long long n, slr_guided_result *__restrict__ out, slr_guided_hit *raw_out,
                    int raw_cap, unsigned long long *vis_all, uint32_t vis_lg_alloc, int max_ed, unsigned long long *work)
{
    __shared__ GWarpShared sh[G_WARPS];
    __shared__ unsigned long long svis[VIS_SMEM ? G_WARPS : 1][VIS_SMEM ? 512 : 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    GWarpShared &W = sh[wib];
    const long long gwarp = (long long)blockIdx.x * G_WARPS + wib;
    unsigned long long *vis;
    if (VIS_SMEM) {
        vis = svis[wib];
        for (int i = lane; i < 512; i += 32) vis[i] = 0ull;
        __syncwarp();
    } else {
        vis = vis_all + ((size_t)gwarp << vis_lg_alloc);
    }
    uint32_t stamp = 0;                                    // the table starts zeroed: stamp 0 = never written
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    const int nchild = 9 * L;

    while (true) {
        unsigned long long qq = 0;
        if (lane == 0) qq = atomicAdd(work, 1ull);
        const long long q = (long long)__shfl_sync(FULL, qq, 0);
        if (q >= n) break;

        const int ed = ed_arr[q], gid = group_id[q], anc = anchor[q];
        const uint2 group = (gid >= 0 && gid < S.n_groups) ? S.groups[gid] : make_uint2(0u, 0u);
        const uint32_t ch = lane < slice_len ? slices[(size_t)q * stride + lane] : 0u;
        W.codes[lane] = (uint8_t)slr_g_code4(ch);
        __syncwarp();
        slr_guided_hit *raw = raw_out ? raw_out + (size_t)q * raw_cap : nullptr;
        const uint32_t vlg = slr_g_vis_log2(ed);
        const bool use_vis = ed >= 2;                      // MINED_TOHASHTESTED_* = 2 (java:L82-L83)
        // candidate filter of the last level (guided_core.cuh): lane i holds slots i and 32 + i of a small group's table
        uint32_t cand0, cand1;
        const bool cok0 = slr_g_filter_slot(S.slots, group, (uint32_t)lane, cand0), cok1 = slr_g_filter_slot(S.slots, group, 32u + lane, cand1);
        // usable when the group is small and, in the BC flavour, the global lists are not probed at the last level (= ed)
        const bool filt_leaf = slr_g_filter_usable(group) &&
                               !(S.bc_flavour && (((S.all_set.y & 0x200u) && ed <= S.all_ed) || ((S.empty_set.y & 0x200u) && ed <= S.empty_ed)));
        // far-node test one level above the last: usable when the global lists are out of reach at the levels ed-1 and ed
        const bool filt_inner = ed >= 2 && slr_g_filter_usable(group) &&
                                !(S.bc_flavour && (((S.all_set.y & 0x200u) && ed - 1 <= S.all_ed) || ((S.empty_set.y & 0x200u) && ed - 1 <= S.empty_ed)));
        if (filt_inner) {
            const SlrGPeq p0 = slr_g_peq(cand0, L), p1 = slr_g_peq(cand1, L);
            for (int b = 0; b < 4; b++) { W.peq[0][b][lane] = p0.eq[b]; W.peq[1][b][lane] = p1.eq[b]; }
        }
        SlrGTop2 T;
        slr_g_top2_init(T);
        uint32_t flags = (ed < 0 || ed > max_ed) ? SLR_G_EXCEPTION : 0u;      // host entry points refuse such a batch

        for (int k = 0; k <= 2 * plusminus && !flags; k++) {
            const int o = slr_g_offset_of(k), ws = anc + o;
            if (ws < 0 || ws + L + post_len > slice_len) { flags = SLR_G_EXCEPTION; break; }           // getSubSequence throws
            const uint32_t c4 = lane < L + post_len ? W.codes[ws + lane] : 0u;
            // N (15) or a char outside ENCODE_MATRIX (-1) in the window indexes outside FOURBIT_TO_TWOBIT_MATRIX (AIOOBE)
            if (__any_sync(FULL, lane < L && c4 >= 15u)) { flags = SLR_G_EXCEPTION; break; }
            const uint32_t postbad = __ballot_sync(FULL, lane >= L && lane < L + post_len && c4 == 0xFFu) >> L;   // thrown only if used
            bool throws = false, throws_full = false;
            const uint32_t two = slr_g_two_of_code4(c4);
            const uint32_t w = __reduce_or_sync(FULL, lane < L ? two << (2 * (L - 1 - lane)) : 0u);     // getLongHashForBytes
            const uint32_t post2 = __reduce_or_sync(FULL, (lane >= L && lane < L + post_len && lane - L < 16) ? two << (2 * (lane - L)) : 0u);
            stamp++;
            int nlist = 0;                                 // matchingList.size() of this tester

            // ---- root: checkMatchWithTestSets(parent) with currentlevel 1 (java:L82-L89) ----
            uint32_t root_meta = slr_g_root_meta();
            {
                bool inh;
                const uint32_t where = slr_g_probe(S, group, w, root_meta, 1, inh);
                if (where) { g_record(T, S.bc_flavour, w, root_meta, 1, o, where, raw, raw_cap, lane); nlist++; }
                if (inh) root_meta |= 1u << 23;
            }
            if (ed == 0) continue;                         // java:L91-L92
            int sp = 0;
            if (lane == 0) { W.stack[0].seq = w; W.stack[0].meta = root_meta; }
            sp = 1;
            __syncwarp();

            while (sp > 0) {
                const SlrGNode node = W.stack[--sp];       // pollLast (java:L100)
                __syncwarp();
                const int level = slr_g_level(node.meta), pos_prev = slr_g_pos_prev(node.meta);
                if (level == ed) {
                    // ---- last level: every position of this node, 32 children per step ----
                    const int p0 = pos_prev == 0 ? 1 : 0;  // first position that runs; node.seq is "tested" after it (java:L122)
                    // skip the batch when no child can hit: a dead node, or (small group, the global lists out of reach at this level, no
                    // invalid post base a deletion would trip over) no candidate within one edit of the node; else run only the
                    // positions at which some candidate can be hit
                    int c_lo = 0, c_hi = nchild - 1;
                    bool run = !slr_g_dead(node.meta);
                    if (run && filt_leaf && postbad == 0u) {
                        int a0 = L, b0 = -1, a1 = L, b1 = -1;
                        if (cok0) slr_g_may_be_child(cand0, node.seq, L, a0, b0);
                        if (cok1) slr_g_may_be_child(cand1, node.seq, L, a1, b1);
                        const int pmin = __reduce_min_sync(FULL, a0 < a1 ? a0 : a1), pmax = __reduce_max_sync(FULL, b0 > b1 ? b0 : b1);
                        run = pmin <= pmax;
                        c_lo = 9 * pmin; c_hi = 9 * pmax + 8;
                    }
                    for (int c0 = c_lo; run && c0 <= c_hi; c0 += 32) {
                        const int c = c0 + lane, p = c / 9, j = c - 9 * p;
                        bool valid = false, inh;
                        uint32_t cmeta = 0, s = 0, where = 0;
                        if (c <= c_hi && p != pos_prev) {                                                // java:L109-L110
                            s = slr_g_child(node.seq, node.meta, L, p, j, post2, postbad, post_len, valid, cmeta, throws);
                            if (valid && use_vis && ((s == node.seq && p > p0) || slr_g_vis_contains(vis, vlg, stamp, s))) valid = false;
                            if (valid) where = slr_g_probe(S, group, s, cmeta, level, inh);
                        }
                        uint32_t hits = __ballot_sync(FULL, where != 0u);
                        while (hits) {
                            const int src = __ffs(hits) - 1;
                            hits &= hits - 1;
                            g_record(T, S.bc_flavour, __shfl_sync(FULL, s, src), __shfl_sync(FULL, cmeta, src), level, o,
                                     __shfl_sync(FULL, where, src), raw, raw_cap, lane);
                            nlist++;
                        }
                    }
                    if (use_vis) {
                        if (lane == 0 && !slr_g_vis_insert(vis, vlg, stamp, node.seq)) throws_full = true;
                        __syncwarp();
                    }
                    continue;
                }
                // ---- a fresh node one level above the last that is far from every candidate (or dead): nothing in its subtree can hit, all
                // that remains of it are the visited-set entries of the node and of its children (each child would be popped, run
                // its — hitless — batch and mark itself) ----
                if (level == ed - 1 && slr_g_pos_cur(node.meta) < 0 && filt_inner && postbad == 0u) {
                    bool far = slr_g_dead(node.meta);
                    if (!far) {
                        SlrGPeq p0, p1;
                        for (int b = 0; b < 4; b++) { p0.eq[b] = W.peq[0][b][lane]; p1.eq[b] = W.peq[1][b][lane]; }
                        far = !__any_sync(FULL, (cok0 && slr_g_within2(p0, node.seq, L)) || (cok1 && slr_g_within2(p1, node.seq, L)));
                    }
                    if (far) {
                        if (bailout < 0 || level < bailout || nlist == 0) {             // else no child is pushed (java:L134-L135)
                            for (int c0 = 0; c0 < nchild; c0 += 32) {
                                const int c = c0 + lane, p = c / 9, j = c - 9 * p;
                                if (c < nchild && p != pos_prev) {
                                    bool valid = false, thr = false;
                                    uint32_t cmeta;
                                    const uint32_t sc = slr_g_child(node.seq, node.meta, L, p, j, post2, 0u, post_len, valid, cmeta, thr);
                                    if (valid && !slr_g_vis_insert_atomic(vis, vlg, stamp, sc)) throws_full = true;
                                }
                            }
                        }
                        if (lane == 0 && !slr_g_vis_insert_atomic(vis, vlg, stamp, node.seq)) throws_full = true;
                        __syncwarp();
                        continue;
                    }
                }
                // ---- inner node: one position (java:L104-L122) ----
                const int pos = slr_g_pos_cur(node.meta) + 1;
                const uint32_t meta = (node.meta & ~31u) | (uint32_t)(pos + 1);
                if (pos < L - 1) {                         // continuation, pushed BEFORE the children (java:L105-L106)
                    if (lane == 0) { W.stack[sp].seq = node.seq; W.stack[sp].meta = meta; }
                    sp++;
                }
                if (pos_prev == pos) { __syncwarp(); continue; }
                bool valid = false, inh = false;
                uint32_t cmeta = 0, s = 0, where = 0;
                if (lane < 9) {
                    s = slr_g_child(node.seq, meta, L, pos, lane, post2, postbad, post_len, valid, cmeta, throws);
                    if (valid && use_vis && slr_g_vis_contains(vis, vlg, stamp, s)) valid = false;
                    if (valid) where = slr_g_probe(S, group, s, cmeta, level, inh);
                }
                uint32_t hits = __ballot_sync(FULL, where != 0u);
                // goNextEDlevel (java:L134-L142): ed > level holds here; bailout: no push once level >= bailout and the list is non-empty
                const bool push_ok = valid && (bailout < 0 || level < bailout || nlist + __popc(hits & le_mask) == 0);
                while (hits) {
                    const int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    g_record(T, S.bc_flavour, __shfl_sync(FULL, s, src), __shfl_sync(FULL, cmeta, src), level, o,
                             __shfl_sync(FULL, where, src), raw, raw_cap, lane);
                    nlist++;
                }
                const uint32_t pm = __ballot_sync(FULL, push_ok);
                if (push_ok) {
                    const int at = sp + __popc(pm & lt_mask);
                    W.stack[at].seq = s;
                    W.stack[at].meta = slr_g_next_level_meta((cmeta & ~(1u << 23)) | ((uint32_t)inh << 23), pos);
                }
                sp += __popc(pm);
                if (use_vis && lane == 0 && !slr_g_vis_insert(vis, vlg, stamp, node.seq)) throws_full = true;   // addToTestedSeqs (java:L122)
                __syncwarp();
            }
            if (__any_sync(FULL, throws)) flags = SLR_G_EXCEPTION;     // a deletion used an invalid post base somewhere in this window
            if (__any_sync(FULL, throws_full)) flags |= SLR_G_TABLE_FULL;
        }
        if (lane == 0) slr_g_top2_store(T, flags, out[q]);
        __syncwarp();
    }
}

}  // namespace

// one wave of persistent CTAs: #SMs x resident CTAs per SM (registers / shared memory decide).  Returns the bytes of per-warp
// visited tables that wave needs in global memory (8 for ed <= 2; 256 KB per warp at ed 3, 1 MB at ed 4).
size_t slr_guided_vis_bytes(int max_ed, int *warps_out)
{
    static int per_sm[64][2], sms_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    const int v = max_ed <= 2 ? 1 : 0;
    if (per_sm[dev][v] == 0) {
        int bps = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (v) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, guided_match_kernel<true>, G_WARPS * 32, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, guided_match_kernel<false>, G_WARPS * 32, 0);
        per_sm[dev][v] = bps > 0 ? bps : 1;
        sms_of[dev] = sms;
    }
    const int sms = sms_of[dev];
    const int ctas = per_sm[dev][v] * sms;
    const int warps = ctas * G_WARPS;
    if (warps_out) *warps_out = warps;
    if (max_ed <= 2) return 8;
    return ((size_t)warps << slr_g_vis_log2(max_ed)) * sizeof(unsigned long long);
}

cudaError_t slr_launch_guided_match(const SlrGuidedSetsDev &S, int L, int plusminus, int post_len, int bailout, const uint8_t *d_slices,
                                    int stride, int slice_len, const int32_t *d_anchor, const int32_t *d_group_id, const int32_t *d_ed,
                                    int max_ed, long long n, slr_guided_result *d_out, slr_guided_hit *d_raw, int raw_cap, void *d_vis,
                                    unsigned long long *d_work, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int warps = 0;
    const size_t vbytes = slr_guided_vis_bytes(max_ed, &warps);
    cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_vis, 0, vbytes, stream);
    if (e != cudaSuccess) return e;
    const long long need = (n + G_WARPS - 1) / G_WARPS;
    const long long resident = warps / G_WARPS;
    const unsigned blocks = (unsigned)(need < resident ? need : resident);
    if (max_ed <= 2)
        guided_match_kernel<true><<<blocks, G_WARPS * 32, 0, stream>>>(S, L, plusminus, post_len, bailout, d_slices, stride, slice_len, d_anchor,
                                                                      d_group_id, d_ed, n, d_out, d_raw, raw_cap, (unsigned long long *)d_vis,
                                                                      slr_g_vis_log2(max_ed), max_ed, d_work);
    else
        guided_match_kernel<false><<<blocks, G_WARPS * 32, 0, stream>>>(S, L, plusminus, post_len, bailout, d_slices, stride, slice_len, d_anchor,
                                                                       d_group_id, d_ed, n, d_out, d_raw, raw_cap, (unsigned long long *)d_vis,
                                                                       slr_g_vis_log2(max_ed), max_ed, d_work);
    return cudaGetLastError();
}
