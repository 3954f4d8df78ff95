// bc_assign.cu — B200 (sm_100a) kernel for the barcode-assignment hot path.
//
// Replaces, for a whole chunk of reads, the body of Parser.assignBarcode
// (F!com/rw/nanoporereadscanner/analyzers/Parser.class, Parser.java:L198-L252): 2*plusminus+1 runs of
// BarcodeMatchTester.doJob (BarcodeMatchTester.java:L198-L244) per read + the Matches merge and the
// best / second-best decision.
//
// The reference enumerates the <=ED edit neighbourhood of each 16-nt window depth-first and probes a hash
// set per mutant, keeping for each ED level only the FIRST hit in its traversal order.  This kernel returns
// bit-identical records but never walks mutant by mutant:
//   * one warp per read, windows in the reference's offset order 0,-1,+1,-2,+2;
//   * a node's single-edit neighbourhood is tested with 12 bucket loads (4 digit groups x {SUB,INS,DEL},
//     slr_table.cuh) instead of ~115 probes; a slot that passes the tag filter is decoded back to the
//     (position, base) pairs that generate it, which gives its rank in the reference's traversal order;
//   * "first hit wins" = warp min-reduction (__reduce_min_sync) over those ranks; level-2 nodes are
//     expanded two per warp step in the reference's LIFO processing order, so the search for the ED-2 slot
//     stops at the first step that yields a valid hit (the reference keeps enumerating to no effect);
//   * the reference's partial visited set (NucTwoBitPerBaseEDtesterBase.java:L105-L120: (int) seq of every
//     node already expanded) is a 256-entry per-warp shared-memory hash  value -> earliest processing time,
//     consulted only for real whitelist hits.
// Per-lane logic: bc_core.cuh (shared with tests/host_sim).
#include "bc_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct WarpShared {
    unsigned long long vh[SLR_VH_SIZE];        // visited hash: (value << 32) | processing time
    uint8_t live[160];                         // level-1 nodes to expand, in processing order
    SlrMatchStore ms;
};

__device__ __forceinline__ void vh_insert(unsigned long long *tab, uint32_t v, uint32_t t)
{
    uint32_t slot = slr_vh_slot(v);
    const unsigned long long val = ((unsigned long long)v << 32) | t;
    while (true) {
        unsigned long long cur = *((volatile unsigned long long *)&tab[slot]);
        if (cur == SLR_VH_EMPTY) {
            const unsigned long long old = atomicCAS(&tab[slot], SLR_VH_EMPTY, val);
            if (old == SLR_VH_EMPTY) return;
            cur = old;
        }
        if ((uint32_t)(cur >> 32) == v) { atomicMin(&tab[slot], val); return; }
        slot = (slot + 1) & (SLR_VH_SIZE - 1);
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
bc_assign_kernel(SlrTableDev tab, int ed_max, int plusminus, int three_prime, const uint8_t *__restrict__ slices,
                 int stride, int slice_len, const int32_t *__restrict__ lens, const int32_t *__restrict__ anchor,
                 long long n, slr_bc_result *__restrict__ out)
{
    __shared__ WarpShared smem[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const long long read = (long long)blockIdx.x * WARPS_PER_BLOCK + wib;
    if (read >= n) return;
    WarpShared &S = smem[wib];

    // ---- the slice: lane i owns char i; bit planes by ballot ------------------------------------------------
    const int len = lens ? min(lens[read], slice_len) : slice_len;
    const uint32_t ch = (lane < len) ? (uint32_t)slices[read * (long long)stride + lane] : 0u;
    const int anc = anchor[read];
    const uint32_t c2 = slr_code2(ch);
    SlrSliceBits sb;
    sb.bit0 = __ballot_sync(FULL, c2 & 1u);
    sb.bit1 = __ballot_sync(FULL, (c2 >> 1) & 1u);
    sb.nonacgt = __ballot_sync(FULL, c2 == 4u);
    sb.unknown = __ballot_sync(FULL, !slr_in_encode_matrix(ch));
    sb.over253 = __ballot_sync(FULL, ch >= 254u);

    uint32_t flags = 0;
    const int noff = 2 * plusminus + 1;
    if (lane < SLR_MAX_OFFSETS) S.ms.m_valid[lane] = 0;
    __syncwarp();

    const int g = (lane >> 2) & 3, op = lane & 3, h = lane >> 4;

    for (int k = 0; k < noff; k++) {
        uint32_t w, p1, p2;
        bool dead_window;
        if (!slr_window(sb, len, anc, slr_offset_of(k), three_prime, ed_max, w, p1, p2, dead_window)) {
            flags |= SLR_F_EXCEPTION;
            break;
        }
        if (lane == 0) S.ms.m_w[k] = w;
        if (dead_window) continue;

        // ======== BarcodeMatchTester.doJob for this window ====================================================
        uint32_t valid_levels = 0;
        if (lane == 3) valid_levels = slr_contains(tab, w) ? 1u : 0u;             // ED 0 (L204-L206)
        valid_levels = __shfl_sync(FULL, valid_levels, 3);
        if (valid_levels && lane == 0) { S.ms.m_bc[k][0] = w; S.ms.m_cnt[k][0] = 0; }

        if (ed_max >= 1) {
            const bool use_vis = ed_max >= 2;
            if (use_vis) {
                // visited hash over the 139 level-1 mutants; processing time t = p*16 + (8-j): the nodes of one
                // root position are popped in reverse creation order (ArrayDeque add / pollLast, L212-L218)
                for (int i = lane; i < SLR_VH_SIZE; i += 32) S.vh[i] = SLR_VH_EMPTY;
                __syncwarp();
                for (int r = 0; r < 5; r++) {
                    const int sl = r * 32 + lane;
                    if (sl < 144) {
                        const int p = sl / 9, j = 8 - (sl - p * 9);
                        bool v, d;
                        const uint32_t mval = slr_gen_mutant(w, p, j, p1, v, d);
                        if (v) vh_insert(S.vh, mval, (uint32_t)(p * 16 + (8 - j)));
                    }
                }
                __syncwarp();
            }
            SlrExpand e;
            e.cs = w; e.w = w; e.pskip = -1; e.cbase = p1; e.tproc = 0; e.level = 1; e.use_visited = use_vis;
            uint32_t r1 = SLR_NONE32, bc1 = 0;
            if (lane < 16 && op < 3) r1 = slr_expand_group(tab, e, S.vh, g, op, bc1);
            const uint32_t rmin = __reduce_min_sync(FULL, r1);
            if (rmin != SLR_NONE32) {                                             // first ED-1 hit in creation order
                const int src = __ffs((int)__ballot_sync(FULL, r1 == rmin)) - 1;
                bc1 = __shfl_sync(FULL, bc1, src);
                valid_levels |= 2u;
                if (lane == 0) { S.ms.m_bc[k][1] = bc1; S.ms.m_cnt[k][1] = (uint8_t)slr_cnt_of(rmin % 9u); }
            }
            if (use_vis) {
                // ---- level-1 nodes the reference expands, in processing order ------------------------------
                int nlive = 0;
                for (int r = 0; r < 5; r++) {
                    const int sl = r * 32 + lane;
                    bool livenode = false;
                    if (sl < 144) {
                        const int p = sl / 9, j = 8 - (sl - p * 9);
                        bool v, d;
                        const uint32_t mval = slr_gen_mutant(w, p, j, p1, v, d);
                        livenode = v && !d && !slr_is_visited(e, S.vh, mval, p);
                    }
                    const uint32_t bal = __ballot_sync(FULL, livenode);
                    if (livenode) S.live[nlive + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)sl;
                    nlive += __popc(bal);
                }
                __syncwarp();
                // ---- level 2: two nodes per warp step; stop at the first step with a valid hit -------------
                for (int it = 0; it * 2 < nlive; it++) {
                    const int idx = it * 2 + h;
                    uint32_t r2 = SLR_NONE32, bc2 = 0, c1 = 0;
                    if (idx < nlive && op < 3) {
                        const int sl = S.live[idx];
                        const int p = sl / 9, jj = sl - p * 9, j = 8 - jj;
                        bool v, d;
                        SlrExpand e2;
                        e2.cs = slr_gen_mutant(w, p, j, p1, v, d);
                        e2.w = w; e2.pskip = p; e2.level = 2; e2.use_visited = true;
                        e2.cbase = (j >= 4 && j < 8) ? p2 : p1;                   // post[nDel+1]; nDel = 1 below an INS node
                        e2.tproc = (uint32_t)(p * 16 + jj);
                        c1 = slr_cnt_of((uint32_t)j);
                        r2 = slr_expand_group(tab, e2, S.vh, g, op, bc2);
                        if (r2 != SLR_NONE32) r2 |= (uint32_t)h << 16;
                    }
                    const uint32_t m2 = __reduce_min_sync(FULL, r2);
                    if (m2 != SLR_NONE32) {
                        const int src = __ffs((int)__ballot_sync(FULL, r2 == m2)) - 1;
                        bc2 = __shfl_sync(FULL, bc2, src);
                        c1 = __shfl_sync(FULL, c1, src);
                        valid_levels |= 4u;
                        if (lane == 0) {
                            S.ms.m_bc[k][2] = bc2;
                            S.ms.m_cnt[k][2] = (uint8_t)(c1 + slr_cnt_of((m2 & 0xFFFFu) % 9u));
                        }
                        break;
                    }
                }
            }
        }
        if (lane == 0) S.ms.m_valid[k] = (uint8_t)valid_levels;
        __syncwarp();
    }
    __syncwarp();

    // ======== merge + decision (Parser.java:L240-L311); uniform across the warp, lane 0 writes ================
    slr_bc_result res;
    res.bc = 0; res.ed = -1; res.ed_second = 0x7FFFFFFF; res.offset = 0; res.n_ins = 0; res.n_del = 0; res.n_sub = 0;
    res.rank = -1; res.flags = flags;
    if (!(flags & SLR_F_EXCEPTION)) {
        const int lv = slr_decide(S.ms, noff, ed_max, res);
        if (lv >= 0 && lane == 0) {
            const int ix = slr_index_of(tab, (uint32_t)res.bc);
            res.rank = (ix >= 0 && tab.rank) ? tab.rank[ix] : ix;                  // CountsRank.rank (L267-L269)
            if (ix >= 0 && tab.counts) atomicAdd(&tab.counts[(size_t)ix * 3 + lv], 1ull);   // BarcodeCounts.addCountForEd (L305-L311)
        }
    }
    if (lane == 0) {
        uint4 *o = reinterpret_cast<uint4 *>(out + read);
        uint4 v0, v1;
        v0.x = (uint32_t)res.bc; v0.y = (uint32_t)(res.bc >> 32); v0.z = (uint32_t)res.ed; v0.w = (uint32_t)res.ed_second;
        v1.x = (uint32_t)(uint8_t)res.offset | ((uint32_t)(uint8_t)res.n_ins << 8) | ((uint32_t)(uint8_t)res.n_del << 16) |
               ((uint32_t)(uint8_t)res.n_sub << 24);
        v1.y = (uint32_t)res.rank; v1.z = res.flags; v1.w = 0;
        o[0] = v0; o[1] = v1;
    }
}

}  // namespace

cudaError_t slr_launch_bc_assign(const SlrTableDev &tab, int ed_max, int plusminus, int three_prime, const uint8_t *d_slices,
                                 int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor, long long n,
                                 slr_bc_result *d_out, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const long long blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    bc_assign_kernel<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(tab, ed_max, plusminus, three_prime, d_slices, stride,
                                                                           slice_len, d_lens, d_anchor, n, d_out);
    return cudaGetLastError();
}
