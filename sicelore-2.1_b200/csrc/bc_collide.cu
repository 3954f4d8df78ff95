// bc_collide.cu — B200 (sm_100a) kernel for the pass-1 barcode collision test.
//
// Replaces the BarcodeDatasetColissionTester.submitSeq loop
// (F!com/rw/nanoporereadscanner/analyzers/BarcodeDatasetColissionTester.class, BarcodeDatasetColissionTester.java:L212-L229):
// one BarcodeMatchTester.doJob per used barcode against the used-barcode list itself, with
//   skipFullMatches = true, allowIndels = true, offset 0, postSeq = null, doNextLevelIfMatchFound = false  (L215-L222).
// Same engine as bc_assign.cu (digit-group buckets, traversal ranks, visited hash, first hit per ED level wins), with the
// three differences those arguments make:
//   * postSeq == null: a deletion appends all four bases (BarcodeMatchTester.java:L336-L340) -> 12 creations per position,
//     and a node needs 21 bucket probes (the appended base is part of `rest` for digit groups 0-2);
//   * doNext == false: a SUB child is expanded only if it HIT, an INS/DEL child only if it did NOT (L268 vs L295, L351), so
//     the visited set and the level-2 frontier depend on the membership of every level-1 mutant (one exact lookup each);
//   * skipFullMatches: a mutant equal to the barcode itself never counts as a hit (L367-L368).
// One warp per barcode; per-lane logic in bc_core.cuh (shared with tests/host_sim).
#include "bc_core.cuh"
#include "slr_kernels.h"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct alignas(16) CollideShared {
    unsigned long long vh[SLR_VH_SIZE];        // (value << 32) | processing time of the EXPANDED level-1 nodes
    uint2 node[192];                           // expanded level-1 nodes in processing order: (sequence, slr_node_meta-like)
};

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

template <int EDMAX>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
bc_collide_kernel(SlrTableDev tab, const unsigned long long *__restrict__ queries, long long n, slr_collide_result *__restrict__ out)
{
    __shared__ CollideShared smem[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    CollideShared &S = smem[wib];
    const long long nwarps = (long long)gridDim.x * WARPS_PER_BLOCK;

    for (long long qi = (long long)blockIdx.x * WARPS_PER_BLOCK + wib; qi < n; qi += nwarps) {
        const unsigned long long q64 = queries[qi];
        uint32_t best1 = SLR_NONE32, bc1 = 0, best2 = SLR_NONE32, bc2b = 0, cnt2 = 0;
        if (!(q64 >> 32) && EDMAX >= 1) {                       // bits >= 32 set: not a clean 16-mer, nothing can match
            const uint32_t w = (uint32_t)q64;
            int nlive = 0;
            if (EDMAX >= 2) {
                ulonglong2 *vh2 = reinterpret_cast<ulonglong2 *>(S.vh);
#pragma unroll
                for (int i = 0; i < SLR_VH_SIZE / 64; i++) vh2[i * 32 + lane] = make_ulonglong2(SLR_VH_EMPTY, SLR_VH_EMPTY);
                __syncwarp();
            }
            // ---- level 1: the 16 x 12 creations in processing order (position ascending, reverse creation order) ----
#pragma unroll 1
            for (int r = 0; r < 6; r++) {
                const int sl = r * 32 + lane;
                const int p = sl / 12, jj = sl - p * 12, j = 11 - jj;
                bool v, d;
                const uint32_t mv = slr_gen_mutant12(w, p & 15, j, v, d);
                const bool member = v && !d && mv != w && slr_contains(tab, mv);   // checkMatchWithTestSets incl. skipFullMatches
                const bool pushtype = v && ((j < 4) == member);                    // goNext: SUB if hit, INS / DEL if not
                bool blocked = false;
                if (EDMAX >= 2) {
                    const uint32_t vmask = __ballot_sync(FULL, pushtype);
                    const uint32_t lowpeers = __match_any_sync(FULL, mv) & vmask & ((1u << lane) - 1u);
                    if (pushtype && lowpeers == 0u) vh_insert_first(S.vh, mv, (uint32_t)(p * 16 + jj));
                    __syncwarp();
                    const uint32_t t = v ? slr_vh_tmin(S.vh, mv) : SLR_NONE32;     // earliest expansion of this value
                    blocked = (p >= 1 && mv == w) || (t != SLR_NONE32 && (int)(t >> 4) < p);
                }
                const bool created = v && !blocked;
                const uint32_t r1 = (created && member) ? (uint32_t)(p * 16 + j) : SLR_NONE32;
                const uint32_t rmin = __reduce_min_sync(FULL, r1);
                if (rmin < best1) {
                    const int src = __ffs((int)__ballot_sync(FULL, r1 == rmin)) - 1;
                    bc1 = __shfl_sync(FULL, mv, src);
                    best1 = rmin;
                }
                if (EDMAX >= 2) {
                    const bool live = created && pushtype && !d;
                    const uint32_t bal = __ballot_sync(FULL, live);
                    if (live) S.node[nlive + __popc(bal & ((1u << lane) - 1u))] =
                                  make_uint2(mv, (uint32_t)(p * 16 + jj) | (slr_cnt_of((uint32_t)j) << 10));
                    nlive += __popc(bal);
                }
            }
            __syncwarp();
            // ---- level 2: 21 probes per expanded node (4 groups x SUB, INS; DEL x 4 appended bases for groups 0-2; DEL group 3) ----
            if (EDMAX >= 2) {
                const int nprobe = nlive * 21;
#pragma unroll 1
                for (int base = 0; base < nprobe; base += 32) {
                    const int pi = base + lane;
                    uint32_t r2 = SLR_NONE32, b2 = 0, c1 = 0;
                    if (pi < nprobe) {
                        const int nd = pi / 21, rem = pi - nd * 21;
                        const int g = rem < 8 ? (rem >> 1) : (rem < 20 ? ((rem - 8) >> 2) : 3);
                        const int op = rem < 8 ? (rem & 1) : 2;
                        const uint2 nrec = S.node[nd];
                        SlrExpand e2 = slr_node_expand(nrec.x, nrec.y, w);
                        e2.nopost = true;
                        e2.cbase = (rem >= 8 && rem < 20) ? (uint32_t)((rem - 8) & 3) : 0u;
                        c1 = nrec.y >> 10;
                        const SlrProbe pr = slr_probe_addr(tab, e2.cs, e2.cbase, g, op);
                        const SlrBucket bk = slr_load_bucket(tab, g, pr.bucket);
                        r2 = slr_probe_eval(tab, e2, S.vh, g, op, pr, bk, b2);
                        if (r2 != SLR_NONE32) r2 |= (uint32_t)nd << 8;
                    }
                    const uint32_t m2 = __reduce_min_sync(FULL, r2);
                    if (m2 < best2) {
                        const int src = __ffs((int)__ballot_sync(FULL, r2 == m2)) - 1;
                        bc2b = __shfl_sync(FULL, b2, src);
                        cnt2 = __shfl_sync(FULL, c1, src);
                        best2 = m2;
                    }
                    if (best2 != SLR_NONE32 && (int)(best2 >> 8) * 21 + 21 <= base + 32) break;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            slr_collide_result res;
            memset(&res, 0, sizeof(res));
            if (best1 != SLR_NONE32) {
                const uint32_t c = slr_cnt_of(best1 & 15u);
                res.valid |= 1u; res.bc[0] = bc1;
                res.n_sub[0] = (uint8_t)(c & 3u); res.n_ins[0] = (uint8_t)((c >> 2) & 3u); res.n_del[0] = (uint8_t)((c >> 4) & 3u);
            }
            if (best2 != SLR_NONE32) {
                const uint32_t c = cnt2 + slr_cnt_of(best2 & 15u);
                res.valid |= 2u; res.bc[1] = bc2b;
                res.n_sub[1] = (uint8_t)(c & 3u); res.n_ins[1] = (uint8_t)((c >> 2) & 3u); res.n_del[1] = (uint8_t)((c >> 4) & 3u);
            }
            out[qi] = res;
        }
        __syncwarp();
    }
}

template <int EDMAX>
cudaError_t launch_t(const SlrTableDev &tab, const unsigned long long *d_queries, long long n, slr_collide_result *d_out, cudaStream_t stream)
{
    int dev = 0, sms = 0, bps = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, bc_collide_kernel<EDMAX>, WARPS_PER_BLOCK * 32, 0);
    if (e != cudaSuccess) return e;
    const long long need = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, resident = (long long)sms * (bps > 0 ? bps : 1);
    bc_collide_kernel<EDMAX><<<(unsigned)(need < resident ? need : resident), WARPS_PER_BLOCK * 32, 0, stream>>>(tab, d_queries, n, d_out);
    return cudaGetLastError();
}

}  // namespace

cudaError_t slr_launch_bc_collide(const SlrTableDev &tab, int ed_max, const unsigned long long *d_queries, long long n,
                                  slr_collide_result *d_out, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    switch (ed_max) {
    case 0: return launch_t<0>(tab, d_queries, n, d_out, stream);
    case 1: return launch_t<1>(tab, d_queries, n, d_out, stream);
    case 2: return launch_t<2>(tab, d_queries, n, d_out, stream);
    default: return cudaErrorInvalidValue;
    }
}
