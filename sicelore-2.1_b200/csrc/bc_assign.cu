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
#include <cstdlib>
#include "bc_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int READ_BATCH = 4;                  // reads a warp takes per visit to the work counter
constexpr unsigned FULL = 0xFFFFFFFFu;
#ifndef SLR_BC_MINB_ED2
#define SLR_BC_MINB_ED2 4                      // resident CTAs per SM the ED-2 kernel is compiled for (register cap 64)
#endif

struct alignas(16) WarpShared {
    unsigned long long vh[SLR_VH_SIZE];        // visited hash: (value << 32) | processing time
    uint2 node[144];                           // level-1 nodes to expand, in processing order: (sequence, slr_node_meta)
    SlrMatchStore ms;
    uint32_t win[SLR_MAX_OFFSETS];             // per window: p1 | p2 << 2 | dead << 4   (the window itself is ms.m_w)
    uint32_t r1[SLR_MAX_OFFSETS];              // per window: traversal rank of the ED-1 hit kept so far
};

// -DSLR_BC_L2HINT=1: the read slices and the records are touched once, evict-first (.cs) accesses (no measurable effect, see slr_table.cuh)
__device__ __forceinline__ uint32_t bc_ld_stream_u8(const uint8_t *p)
{
#if !defined(SLR_BC_L2HINT) || SLR_BC_L2HINT == 0
    return (uint32_t)*p;
#else
    uint32_t v;
    asm("ld.global.cs.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ void bc_st_stream(uint4 *p, uint4 v)
{
#if !defined(SLR_BC_L2HINT) || SLR_BC_L2HINT == 0
    *p = v;
#else
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}

// First insertion wins: the caller inserts in increasing processing time (rounds in order, duplicates inside a
// round removed with __match_any_sync), so an existing key always carries the smaller time.  Returns the time
// stored for v.
__device__ __forceinline__ uint32_t vh_insert_first(unsigned long long *tab, uint32_t v, uint32_t t)
{
    uint32_t slot = slr_vh_slot(v);
    const unsigned long long val = ((unsigned long long)v << 32) | t;
    while (true) {
        unsigned long long cur = *((volatile unsigned long long *)&tab[slot]);
        if (cur == SLR_VH_EMPTY) {
            cur = atomicCAS(&tab[slot], SLR_VH_EMPTY, val);
            if (cur == SLR_VH_EMPTY) return t;
        }
        if ((uint32_t)(cur >> 32) == v) return (uint32_t)cur;
        slot = (slot + 1) & (SLR_VH_SIZE - 1);
    }
}

// Warp-parallel forms of slr_level2_plan / slr_decide (bc_core.cuh, the serial forms the CPU replay uses): lane e = 3 k + lv
// owns entry (window k, ED level lv) of the match store.  With at most 8 entries the merged HashMap stays at capacity 16,
// the sort key of an entry is (ED, offset != 0, bucket, insertion order) and everything is a warp min-reduction; more
// entries (HashMap resizes, possible treeified bin) take the serial path.
struct WarpEntry {
    uint32_t key;       // sort key at capacity 16, SLR_NONE32 = no entry
    uint32_t bc;
    int lv;
};
__device__ __forceinline__ WarpEntry warp_entry(const SlrMatchStore &M, int noff, int lane, int max_lv)
{
    WarpEntry e;
    e.key = SLR_NONE32; e.bc = 0;
    const int k = lane / 3;
    e.lv = lane - 3 * k;
    if (k < noff && e.lv <= max_lv && ((M.m_valid[k] >> e.lv) & 1)) {
        const uint32_t w = M.m_w[k], sp = w ^ (w >> 16);
        e.key = ((uint32_t)e.lv << 20) | ((k != 0 ? 1u : 0u) << 16) | ((sp & 15u) << 8) | (uint32_t)lane;
        e.bc = M.m_bc[k][e.lv];
    }
    return e;
}
__device__ __forceinline__ int warp_level2_plan(const SlrMatchStore &M, int noff, int lane, uint32_t &bcA)
{
    const WarpEntry e = warp_entry(M, noff, lane, 1);
    const int n01 = __popc(__ballot_sync(FULL, e.key != SLR_NONE32));
    bcA = 0;
    if (n01 + noff > 8) return SLR_L2_ALL;
    if (n01 == 0) return SLR_L2_TWO;
    const uint32_t best = __reduce_min_sync(FULL, e.key);
    bcA = __shfl_sync(FULL, e.bc, (int)(best & 0xFFu));
    return __ballot_sync(FULL, e.key != SLR_NONE32 && e.bc != bcA) ? SLR_L2_NONE : SLR_L2_UNTIL;
}
__device__ __forceinline__ int warp_decide(const SlrMatchStore &M, int noff, int ed_max, int lane, slr_bc_result &res)
{
    const WarpEntry e = warp_entry(M, noff, lane, 2);
    const int cnt = __popc(__ballot_sync(FULL, e.key != SLR_NONE32));
    if (cnt > 8) return slr_decide(M, noff, ed_max, res);
    if (cnt == 0) return -1;
    const uint32_t best = __reduce_min_sync(FULL, e.key);
    const int be = (int)(best & 0xFFu), bk = be / 3, blv = (int)(best >> 20);
    const uint32_t bbc = __shfl_sync(FULL, e.bc, be);
    const uint32_t second = __reduce_min_sync(FULL, (e.key != SLR_NONE32 && e.bc != bbc) ? (uint32_t)e.lv : 0x7FFFFFFFu);
    res.ed = blv;
    res.ed_second = (int32_t)second;
    if (blv <= ed_max && blv < (int)second) {                                   // L251-L252
        res.flags |= SLR_F_ASSIGNED;
        res.bc = bbc;
        res.offset = (int8_t)slr_offset_of(bk);
        const uint32_t c = M.m_cnt[bk][blv];
        res.n_sub = (int8_t)(c & 3u);
        res.n_ins = (int8_t)((c >> 2) & 3u);
        res.n_del = (int8_t)((c >> 4) & 3u);
        return blv;
    }
    return -1;
}

// Persistent warps: warp i of the grid takes reads i, i + #warps, ... (reads are i.i.d., a static stride balances).
// Per read:
//   1. the 2*plusminus+1 windows (bit-field extracts of the ballot planes);
//   2. levels 0 and 1 of ALL windows: 12 bucket probes per window (4 digit groups x SUB/INS/DEL), 32 probes per warp step;
//      "first hit wins" = minimum traversal rank over the probes of a window;
//   3. slr_level2_plan: which ED-2 searches can still change the record (most reads: none, or just enough to settle
//      ed_second) - the reference runs all of them, but its HashSet.add / distinctByKey make the rest dead work;
//   4. per remaining window: visited hash + live level-1 nodes, then 32 probes per warp step over the nodes in the
//      reference's LIFO processing order, stopping at the first step whose hit cannot be beaten;
//   5. merge + decision (slr_decide), rank lookup, counter, one 32-byte record.
template <int EDMAX>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, EDMAX >= 2 ? SLR_BC_MINB_ED2 : 4)
bc_assign_kernel(SlrTableDev tab, int plusminus, int three_prime, int need_post, const uint8_t *__restrict__ slices, int stride, int slice_len,
                 const int32_t *__restrict__ lens, const int32_t *__restrict__ anchor, long long n, slr_bc_result *__restrict__ out,
                 unsigned long long *__restrict__ work)
{
    __shared__ WarpShared smem[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpShared &S = smem[wib];
    const long long nwarps = (long long)gridDim.x * WARPS_PER_BLOCK;
    const int noff = 2 * plusminus + 1;

    // Work distribution: the cost of a read varies by an order of magnitude (none / one / all ED-2 searches), so a static stride
    // leaves a long tail; warps take batches of READ_BATCH reads from a device counter instead (first batch: static).
    long long batch0 = ((long long)blockIdx.x * WARPS_PER_BLOCK + wib) * READ_BATCH;
    for (;;) {
    if (batch0 >= n) break;
    const long long batch1 = batch0 + READ_BATCH < n ? batch0 + READ_BATCH : n;
    for (long long read = batch0; read < batch1; read++) {
        // ---- the slice: lane i owns char i; bit planes by ballot --------------------------------------------
        const int len = lens ? min(lens[read], slice_len) : slice_len;
        const uint32_t ch = (lane < len) ? bc_ld_stream_u8(slices + read * (long long)stride + lane) : 0u;
        const int anc = anchor[read];
        const uint32_t c2 = slr_code2(ch);
        SlrSliceBits sb;
        sb.bit0 = __ballot_sync(FULL, c2 & 1u);
        sb.bit1 = __ballot_sync(FULL, (c2 >> 1) & 1u);
        sb.nonacgt = __ballot_sync(FULL, c2 == 4u);
        sb.unknown = __ballot_sync(FULL, !slr_in_encode_matrix(ch));
        sb.over253 = __ballot_sync(FULL, ch >= 254u);

        // ---- 1. windows: lane k computes window k (Parser.java:L205-L221) ---------------------------------------
        uint32_t flags = 0;
        {
            uint32_t w = 0, p1 = 0, p2 = 0;
            bool dead_window = false, ok = true;
            if (lane < noff) ok = slr_window(sb, len, anc, slr_offset_of(lane), three_prime, EDMAX, w, p1, p2, dead_window, need_post != 0);
            if (lane < noff) {
                S.ms.m_w[lane] = w;
                S.ms.m_valid[lane] = 0;
                S.win[lane] = p1 | (p2 << 2) | ((dead_window ? 1u : 0u) << 4);
                S.r1[lane] = SLR_NONE32;
            }
            if (__ballot_sync(FULL, !ok)) flags |= SLR_F_EXCEPTION;        // the Java throws at the first bad window: no record
        }
        __syncwarp();

        if (!(flags & SLR_F_EXCEPTION)) {
            // ---- 2. levels 0 and 1 of every window (BarcodeMatchTester.java:L204-L206 + the root's expansion) ----
            // probe pi = 12 k + (digit group, op) of window k, 32 probes per warp step (5 windows: 2 steps); probe 0 of a window
            // (table 0, rest of w) also answers the ED-0 lookup.  First hit per window = minimum traversal rank over its probes,
            // kept across steps in S.r1.
            constexpr int PER = EDMAX >= 1 ? 12 : 1;
            const int nprobe1 = noff * PER;
#pragma unroll 1
            for (int base = 0; base < nprobe1; base += 32) {
                const int pi = base + lane;
                const int k = pi / PER, rem = pi - k * PER;
                uint32_t r2 = SLR_NONE32, bc2 = 0;
                if (pi < nprobe1) {
                    const uint32_t w = S.ms.m_w[k], wi = S.win[k];
                    if (!(wi & 16u)) {
                        const int g = (rem * 11) >> 5, op = rem - 3 * g;
                        const SlrExpand e1 = slr_root_expand(w, wi & 3u, EDMAX >= 2);
                        const SlrProbe pr = slr_probe_addr(tab, w, wi & 3u, g, op);
                        const SlrBucket bk = slr_load_bucket(tab, g, pr.bucket);
                        if (rem == 0 && slr_contains_in(tab, bk, pr.bucket, pr.tag, w >> 24)) {
                            S.ms.m_bc[k][0] = w; S.ms.m_cnt[k][0] = 0; S.ms.m_valid[k] = 1;
                        }
                        if (EDMAX >= 1) r2 = slr_probe_eval(tab, e1, S.vh, g, op, pr, bk, bc2);
                    }
                }
                if (EDMAX >= 1) {
                    const int k_hi = min(noff - 1, (base + 31) / PER);
#pragma unroll 1
                    for (int kk = base / PER; kk <= k_hi; kk++) {
                        const uint32_t mine = (k == kk) ? r2 : SLR_NONE32;
                        const uint32_t m = __reduce_min_sync(FULL, mine);
                        if (m != SLR_NONE32) {
                            const int src = __ffs((int)__ballot_sync(FULL, mine == m)) - 1;
                            const uint32_t b1 = __shfl_sync(FULL, bc2, src);
                            if (lane == 0 && m < S.r1[kk]) { S.r1[kk] = m; S.ms.m_bc[kk][1] = b1; S.ms.m_cnt[kk][1] = (uint8_t)slr_cnt_of(m & 15u); }
                        }
                    }
                }
            }
            __syncwarp();
            if (EDMAX >= 1 && lane < noff && S.r1[lane] != SLR_NONE32) S.ms.m_valid[lane] |= 2u;
            __syncwarp();

            if (EDMAX >= 2) {
                // ---- 3. which ED-2 searches matter? ----------------------------------------------------------------
                uint32_t bcA = 0;
                const int plan = warp_level2_plan(S.ms, noff, lane, bcA);
                bool have_first = false;                                         // SLR_L2_TWO: an ED-2 hit has been seen (its barcode: bcA)
                // ---- 4. ED-2 searches ---------------------------------------------------------------------------------
#pragma unroll 1
                for (int k = 0; k < noff && plan != SLR_L2_NONE; k++) {
                    const uint32_t wi = S.win[k];
                    if (wi & 16u) continue;
                    const uint32_t w = S.ms.m_w[k], p1 = wi & 3u, p2 = (wi >> 2) & 3u;
                    // One pass over the 144 level-1 slots in processing order (5 rounds of 32): build the visited hash
                    // value -> earliest processing time t = p*16 + (8-j), and keep the nodes the reference expands
                    // (valid, no 62-63 garbage, not "already tested" when created) as compact (sequence, meta) records.
                    int nlive = 0;
                    ulonglong2 *vh2 = reinterpret_cast<ulonglong2 *>(S.vh);
#pragma unroll
                    for (int i = 0; i < SLR_VH_SIZE / 64; i++) vh2[i * 32 + lane] = make_ulonglong2(SLR_VH_EMPTY, SLR_VH_EMPTY);
                    __syncwarp();
#pragma unroll 1
                    for (int r = 0; r < 5; r++) {
                        const int sl = r * 32 + lane;
                        const int p = sl / 9, jj = sl - p * 9, j = 8 - jj;
                        bool v, d;
                        const uint32_t mv = slr_gen_mutant(w, p & 15, j, p1, v, d);
                        v = v && sl < 144;
                        const uint32_t vmask = __ballot_sync(FULL, v);
                        const uint32_t lowpeers = __match_any_sync(FULL, mv) & vmask & ((1u << lane) - 1u);
                        uint32_t tfirst = (uint32_t)(p * 16 + jj);
                        if (v && lowpeers == 0u) tfirst = vh_insert_first(S.vh, mv, tfirst);
                        __syncwarp();
                        tfirst = __shfl_sync(FULL, tfirst, (v && lowpeers != 0u) ? __ffs((int)lowpeers) - 1 : lane);
                        const bool livenode = v && !d && !((p >= 1 && mv == w) || (int)(tfirst >> 4) < p);
                        const uint32_t bal = __ballot_sync(FULL, livenode);
                        if (livenode) S.node[nlive + __popc(bal & ((1u << lane) - 1u))] = make_uint2(mv, slr_node_meta(p, j, p1, p2));
                        nlive += __popc(bal);
                    }
                    __syncwarp();

                    // probe pi = node * 12 + (digit group, op) over the live nodes; 32 probes per warp step; first hit wins =
                    // warp minimum of (node, traversal rank).  The search stops once a hit is known and every probe of its
                    // node has been evaluated (later nodes only have larger ranks).
                    const int nprobe = nlive * 12;
                    uint32_t best = SLR_NONE32, bcb = 0, cntb = 0;
#pragma unroll 1
                    for (int base = 0; base < nprobe; base += 32) {
                        const int pi = base + lane;
                        uint32_t r2 = SLR_NONE32, bc2 = 0, c1 = 0;
                        if (pi < nprobe) {
                            const int nd = pi / 12, rem = pi - nd * 12;
                            const int g = (rem * 11) >> 5, op = rem - 3 * g;
                            const uint2 nrec = S.node[nd];
                            const SlrExpand e2 = slr_node_expand(nrec.x, nrec.y, w);
                            c1 = nrec.y >> 10;
                            const SlrProbe pr = slr_probe_addr(tab, e2.cs, e2.cbase, g, op);
                            const SlrBucket bk = slr_load_bucket(tab, g, pr.bucket);
                            r2 = slr_probe_eval(tab, e2, S.vh, g, op, pr, bk, bc2);
                            if (r2 != SLR_NONE32) r2 |= (uint32_t)nd << 8;
                        }
                        const uint32_t m2 = __reduce_min_sync(FULL, r2);
                        if (m2 < best) {
                            const int src = __ffs((int)__ballot_sync(FULL, r2 == m2)) - 1;
                            bcb = __shfl_sync(FULL, bc2, src);
                            cntb = __shfl_sync(FULL, c1, src);
                            best = m2;
                        }
                        if (best != SLR_NONE32 && (int)(best >> 8) * 12 + 12 <= base + 32) break;
                    }
                    if (best != SLR_NONE32) {
                        if (lane == 0) {
                            S.ms.m_bc[k][2] = bcb;
                            S.ms.m_cnt[k][2] = (uint8_t)(cntb + slr_cnt_of(best & 15u));
                            S.ms.m_valid[k] |= 4u;
                        }
                        if (plan == SLR_L2_UNTIL && bcb != bcA) break;               // ed_second is settled
                        if (plan == SLR_L2_TWO) {
                            if (have_first && bcb != bcA) break;                     // two barcodes at ED 2: unassigned, ed = ed_second = 2
                            if (!have_first) { have_first = true; bcA = bcb; }
                        }
                    }
                    __syncwarp();
                }
                __syncwarp();
            }
        }

        // ======== 5. merge + decision (Parser.java:L240-L311); uniform across the warp, lane 0 writes ============
        slr_bc_result res;
        res.bc = 0; res.ed = -1; res.ed_second = 0x7FFFFFFF; res.offset = 0; res.n_ins = 0; res.n_del = 0; res.n_sub = 0;
        res.rank = -1; res.flags = flags;
        if (!(flags & SLR_F_EXCEPTION)) {
            const int lv = warp_decide(S.ms, noff, EDMAX, lane, res);
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
            bc_st_stream(o, v0); bc_st_stream(o + 1, v1);
        }
        __syncwarp();
    }
    unsigned long long nb = 0;
    if (lane == 0) nb = atomicAdd(work, (unsigned long long)READ_BATCH);
    batch0 = (long long)__shfl_sync(FULL, nb, 0) + nwarps * READ_BATCH;
    }
}

template <int EDMAX>
cudaError_t launch_t(const SlrTableDev &tab, int plusminus, int three_prime, int need_post, const uint8_t *d_slices, int stride, int slice_len,
                     const int32_t *d_lens, const int32_t *d_anchor, long long n, slr_bc_result *d_out, unsigned long long *d_work,
                     cudaStream_t stream)
{
    static int resident_ctas[64];                                // per device: #SMs x resident CTAs per SM
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (resident_ctas[dev] == 0) {
        int bps = 0, sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_assign_kernel<EDMAX>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                       (int)((sizeof(WarpShared) * WARPS_PER_BLOCK * 4 * 100) / (228 * 1024)) + 8);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, bc_assign_kernel<EDMAX>, WARPS_PER_BLOCK * 32, 0);
        if (e != cudaSuccess) return e;
        resident_ctas[dev] = sms * (bps > 0 ? bps : 1);
    }
    const long long need = (n + WARPS_PER_BLOCK * READ_BATCH - 1) / (WARPS_PER_BLOCK * READ_BATCH);
    const long long resident = resident_ctas[dev];               // one wave of persistent CTAs: a multiple of the SM count
    const unsigned blocks = (unsigned)(need < resident ? need : resident);
    // SLR_BC_APW=<hit ratio> (experiment, off by default): an access-policy window over the bucket tables for this launch (persisting hits,
    // streaming misses).  Measured: 1.0 -> 48.08 ms vs 48.06 ms without, 0.6 -> 50.97 ms: the table (64 MiB) does not fit one die's half of the L2
    // next to the index map and the counters whatever the policy
    static const float apw = [] { const char *e = getenv("SLR_BC_APW"); return e ? (float)atof(e) : 0.0f; }();
    if (apw > 0.0f) {
        static bool limit_set[64];
        if (!limit_set[dev]) {
            int maxp = 0;
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
            limit_set[dev] = true;
        }
        int maxw = 0;
        cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeAccessPolicyWindow;
        size_t bytes = (size_t)128 << tab.bbits;
        if (maxw > 0 && bytes > (size_t)maxw) bytes = (size_t)maxw;
        at[0].val.accessPolicyWindow.base_ptr = const_cast<uint4 *>(tab.bk);
        at[0].val.accessPolicyWindow.num_bytes = bytes;
        at[0].val.accessPolicyWindow.hitRatio = apw > 1.0f ? 1.0f : apw;
        at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(WARPS_PER_BLOCK * 32); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, bc_assign_kernel<EDMAX>, tab, plusminus, three_prime, need_post, d_slices, stride, slice_len, d_lens, d_anchor,
                                  n, d_out, d_work);
    }
    bc_assign_kernel<EDMAX><<<blocks, WARPS_PER_BLOCK * 32, 0, stream>>>(tab, plusminus, three_prime, need_post, d_slices, stride, slice_len, d_lens,
                                                                        d_anchor, n, d_out, d_work);
    return cudaGetLastError();
}

}  // namespace

cudaError_t slr_launch_bc_assign(const SlrTableDev &tab, int ed_max, int plusminus, int three_prime, int need_post, const uint8_t *d_slices,
                                 int stride, int slice_len, const int32_t *d_lens, const int32_t *d_anchor, long long n,
                                 slr_bc_result *d_out, unsigned long long *d_work, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    cudaError_t e0 = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);      // the work counter of this launch
    if (e0 != cudaSuccess) return e0;
    switch (ed_max) {
    case 0: return launch_t<0>(tab, plusminus, three_prime, need_post, d_slices, stride, slice_len, d_lens, d_anchor, n, d_out, d_work, stream);
    case 1: return launch_t<1>(tab, plusminus, three_prime, need_post, d_slices, stride, slice_len, d_lens, d_anchor, n, d_out, d_work, stream);
    case 2: return launch_t<2>(tab, plusminus, three_prime, need_post, d_slices, stride, slice_len, d_lens, d_anchor, n, d_out, d_work, stream);
    default: return cudaErrorInvalidValue;
    }
}
