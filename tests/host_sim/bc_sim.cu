// bc_sim.cu — TEST INFRASTRUCTURE: replays the warp orchestration of bc_assign.cu sequentially on the CPU,
// using the very same __host__ __device__ per-lane code (bc_core.cuh, slr_table.cuh) and the same table
// builder.  Lets the algorithm (digit-group buckets, traversal ranks, visited-time logic, HashSet order) be
// checked against the oracle in the GPU-less build container.  Never part of the product library.
#include <cstdio>
#include <vector>
#include "../../sicelore-2.1_b200/csrc/bc_core.cuh"
#include "../../sicelore-2.1_b200/csrc/slr_table_build.h"

// first insertion wins; the caller inserts in increasing processing time (like the kernel's rounds)
static void vh_insert_host(unsigned long long *tab, uint32_t v, uint32_t t)
{
    uint32_t slot = slr_vh_slot(v);
    const unsigned long long val = ((unsigned long long)v << 32) | t;
    while (true) {
        unsigned long long cur = tab[slot];
        if (cur == SLR_VH_EMPTY) { tab[slot] = val; return; }
        if ((uint32_t)(cur >> 32) == v) return;
        slot = (slot + 1) & (SLR_VH_SIZE - 1);
    }
}

static int sim_need_post = 1;  // sim_set_need_post(0): the pass-1 exact lookup (window only)
extern "C" void sim_set_need_post(int v) { sim_need_post = v; }
static int force_all_l2 = 0;   // sim_set_force_all_l2(1): run every ED-2 search like the reference (checks that skipping is exact)
extern "C" void sim_set_force_all_l2(int v) { force_all_l2 = v; }

static void sim_read(const SlrTableDev &tab, int ed_max, int plusminus, int three_prime, const uint8_t *slice, int len, int anc,
                     slr_bc_result *out, long long *n_loads)
{
    SlrSliceBits sb = {0, 0, 0, 0, 0};
    for (int lane = 0; lane < 32; lane++) {
        const uint32_t ch = lane < len ? slice[lane] : 0u;
        const uint32_t c2 = slr_code2(ch);
        sb.bit0 |= (c2 & 1u) << lane;
        sb.bit1 |= ((c2 >> 1) & 1u) << lane;
        sb.nonacgt |= (uint32_t)(c2 == 4u) << lane;
        sb.unknown |= (uint32_t)(!slr_in_encode_matrix(ch)) << lane;
        sb.over253 |= (uint32_t)(ch >= 254u) << lane;
    }
    static thread_local unsigned long long vh[SLR_VH_SIZE];
    uint32_t node_cs[144], node_meta[144];
    SlrMatchStore ms;
    memset(&ms, 0, sizeof(ms));
    uint32_t flags = 0;
    const int noff = 2 * plusminus + 1;
    uint32_t win_p1[SLR_MAX_OFFSETS], win_p2[SLR_MAX_OFFSETS];
    bool win_dead[SLR_MAX_OFFSETS];
    // 1. windows
    for (int k = 0; k < noff; k++) {
        uint32_t w = 0, p1 = 0, p2 = 0;
        bool dead_window = false;
        if (!slr_window(sb, len, anc, slr_offset_of(k), three_prime, ed_max, w, p1, p2, dead_window, sim_need_post != 0)) flags |= SLR_F_EXCEPTION;
        ms.m_w[k] = w; win_p1[k] = p1; win_p2[k] = p2; win_dead[k] = dead_window;
    }
    // 2. levels 0 and 1 of every window
    for (int k = 0; k < noff && !(flags & SLR_F_EXCEPTION); k++) {
        if (win_dead[k]) continue;
        const uint32_t w = ms.m_w[k], p1 = win_p1[k];
        const SlrExpand e = slr_root_expand(w, p1, ed_max >= 2);
        uint32_t valid_levels = 0;
        uint32_t rmin = SLR_NONE32, bc1 = 0;
        for (int lane = 0; lane < (ed_max >= 1 ? 12 : 1); lane++) {
            const int g = (lane * 11) >> 5, op = lane - 3 * g;
            const SlrProbe pr = slr_probe_addr(tab, w, p1, g, op);
            const SlrBucket bk = slr_load_bucket(tab, g, pr.bucket);
            (*n_loads)++;
            if (lane == 0 && slr_contains_in(tab, bk, pr.bucket, pr.tag, w >> 24)) valid_levels = 1u;
            if (ed_max >= 1) {
                uint32_t b = 0;
                const uint32_t r = slr_probe_eval(tab, e, vh, g, op, pr, bk, b);
                if (r < rmin) { rmin = r; bc1 = b; }
            }
        }
        if (valid_levels) { ms.m_bc[k][0] = w; ms.m_cnt[k][0] = 0; }
        if (rmin != SLR_NONE32) { valid_levels |= 2u; ms.m_bc[k][1] = bc1; ms.m_cnt[k][1] = (uint8_t)slr_cnt_of(rmin & 15u); }
        ms.m_valid[k] = (uint8_t)valid_levels;
    }
    // 3. + 4. the ED-2 searches that can still change the record
    uint32_t bcA = 0;
    bool have_first = false;
    const int plan = (ed_max >= 2 && !(flags & SLR_F_EXCEPTION)) ? (force_all_l2 ? (int)SLR_L2_ALL : slr_level2_plan(ms, noff, bcA)) : (int)SLR_L2_NONE;
    for (int k = 0; k < noff && plan != SLR_L2_NONE; k++) {
        if (win_dead[k]) continue;
        const uint32_t w = ms.m_w[k], p1 = win_p1[k], p2 = win_p2[k];
        for (int i = 0; i < SLR_VH_SIZE; i++) vh[i] = SLR_VH_EMPTY;
        for (int sl = 0; sl < 144; sl++) {
            const int p = sl / 9, j = 8 - (sl - p * 9);
            bool v, d;
            const uint32_t mval = slr_gen_mutant(w, p, j, p1, v, d);
            if (v) vh_insert_host(vh, mval, (uint32_t)(p * 16 + (8 - j)));
        }
        int nlive = 0;
        for (int sl = 0; sl < 144; sl++) {
            const int p = sl / 9, j = 8 - (sl - p * 9);
            bool v, d;
            const uint32_t mval = slr_gen_mutant(w, p, j, p1, v, d);
            const uint32_t tfirst = v ? slr_vh_tmin(vh, mval) : 0u;
            if (v && !d && !((p >= 1 && mval == w) || (int)(tfirst >> 4) < p)) {
                node_cs[nlive] = mval;
                node_meta[nlive++] = slr_node_meta(p, j, p1, p2);
            }
        }
        const int nprobe = nlive * 12;
        uint32_t best = SLR_NONE32, bcb = 0, cntb = 0;
        for (int base = 0; base < nprobe; base += 32) {
            for (int lane = 0; lane < 32; lane++) {
                const int pi = base + lane;
                if (pi >= nprobe) break;
                const int nd = pi / 12, rem = pi - nd * 12;
                const int g = (rem * 11) >> 5, op = rem - 3 * g;
                const SlrExpand e2 = slr_node_expand(node_cs[nd], node_meta[nd], w);
                uint32_t b = 0;
                uint32_t r2 = slr_expand_group(tab, e2, vh, g, op, b);
                (*n_loads)++;
                if (r2 != SLR_NONE32) r2 |= (uint32_t)nd << 8;
                if (r2 < best) { best = r2; bcb = b; cntb = node_meta[nd] >> 10; }
            }
            if (best != SLR_NONE32 && (int)(best >> 8) * 12 + 12 <= base + 32) break;
        }
        if (best != SLR_NONE32) {
            ms.m_valid[k] |= 4u;
            ms.m_bc[k][2] = bcb;
            ms.m_cnt[k][2] = (uint8_t)(cntb + slr_cnt_of(best & 15u));
            if (plan == SLR_L2_UNTIL && bcb != bcA) break;
            if (plan == SLR_L2_TWO) {
                if (have_first && bcb != bcA) break;
                if (!have_first) { have_first = true; bcA = bcb; }
            }
        }
    }
    slr_bc_result res;
    memset(&res, 0, sizeof(res));
    res.ed = -1; res.ed_second = 0x7FFFFFFF; res.rank = -1; res.flags = flags;
    if (!(flags & SLR_F_EXCEPTION)) {
        const int lv = slr_decide(ms, noff, ed_max, res);
        if (lv >= 0) {
            const int ix = slr_index_of(tab, (uint32_t)res.bc);
            res.rank = (ix >= 0 && tab.rank) ? tab.rank[ix] : ix;
            if (ix >= 0 && tab.counts) tab.counts[(size_t)ix * 3 + lv]++;
        }
    }
    *out = res;
}

extern "C" int sim_bc_assign(const uint64_t *keys, const int32_t *rank, int64_t n_keys, int force_bbits, int ed_max, int plusminus,
                             int three_prime, const uint8_t *slices, int stride, int slice_len, const int32_t *lens,
                             const int32_t *anchor, int64_t n, slr_bc_result *out, unsigned long long *counts, long long *n_loads,
                             long long *stash_sizes)
{
    SlrTableHost T;
    slr_build_table(keys, rank, n_keys, T, force_bbits);
    SlrTableDev tab = slr_table_host_view(T, counts);
    long long loads = 0;
    for (int64_t i = 0; i < n; i++) {
        const int len = lens ? (lens[i] < slice_len ? lens[i] : slice_len) : slice_len;
        sim_read(tab, ed_max, plusminus, three_prime, slices + i * stride, len, anchor[i], &out[i], &loads);
    }
    if (n_loads) *n_loads = loads;
    if (stash_sizes) for (int g = 0; g < 4; g++) stash_sizes[g] = T.st_n[g];
    return 0;
}

// ---- pass-1 collision tester: the orchestration of bc_collide.cu, lane by lane -----------------------------------
extern "C" int sim_bc_collide(const uint64_t *keys, int64_t n_keys, int force_bbits, int ed_max, const uint64_t *queries, int64_t n,
                              slr_collide_result *out, long long *n_loads)
{
    SlrTableHost T;
    slr_build_table(keys, nullptr, n_keys, T, force_bbits);
    SlrTableDev tab = slr_table_host_view(T, nullptr);
    static thread_local unsigned long long vh[SLR_VH_SIZE];
    long long loads = 0;
    for (int64_t qi = 0; qi < n; qi++) {
        slr_collide_result res;
        memset(&res, 0, sizeof(res));
        uint32_t best1 = SLR_NONE32, bc1 = 0, best2 = SLR_NONE32, bc2 = 0, cnt2 = 0;
        if (!(queries[qi] >> 32) && ed_max >= 1) {
            const uint32_t w = (uint32_t)queries[qi];
            uint32_t node_cs[192], node_meta[192];
            int nlive = 0;
            for (int i = 0; i < SLR_VH_SIZE; i++) vh[i] = SLR_VH_EMPTY;
            for (int sl = 0; sl < 192; sl++) {                       // processing order; the kernel does 32 slots per round:
                const int p = sl / 12, jj = sl - p * 12, j = 11 - jj;  // insertions of a round are visible to its own lookups, and a
                bool v, d;                                           // later slot never has a smaller time, so slot order is equivalent
                const uint32_t mv = slr_gen_mutant12(w, p, j, v, d);
                loads += (v && !d && mv != w) ? 1 : 0;
                const bool member = v && !d && mv != w && slr_contains(tab, mv);
                const bool pushtype = v && ((j < 4) == member);
                bool blocked = false;
                if (ed_max >= 2) {
                    // the kernel inserts the whole round first: an insertion of a LATER slot of the same round has a larger time
                    // and the same or a later position, so it cannot block this slot either
                    if (pushtype) vh_insert_host(vh, mv, (uint32_t)(p * 16 + jj));
                    const uint32_t t = v ? slr_vh_tmin(vh, mv) : SLR_NONE32;
                    blocked = (p >= 1 && mv == w) || (t != SLR_NONE32 && (int)(t >> 4) < p);
                }
                const bool created = v && !blocked;
                if (created && member && (uint32_t)(p * 16 + j) < best1) { best1 = (uint32_t)(p * 16 + j); bc1 = mv; }
                if (ed_max >= 2 && created && pushtype && !d) {
                    node_cs[nlive] = mv;
                    node_meta[nlive++] = (uint32_t)(p * 16 + jj) | (slr_cnt_of((uint32_t)j) << 10);
                }
            }
            if (ed_max >= 2) {
                const int nprobe = nlive * 21;
                for (int base = 0; base < nprobe; base += 32) {
                    for (int lane = 0; lane < 32; lane++) {
                        const int pi = base + lane;
                        if (pi >= nprobe) break;
                        const int nd = pi / 21, rem = pi - nd * 21;
                        const int g = rem < 8 ? (rem >> 1) : (rem < 20 ? ((rem - 8) >> 2) : 3);
                        const int op = rem < 8 ? (rem & 1) : 2;
                        SlrExpand e2 = slr_node_expand(node_cs[nd], node_meta[nd], w);
                        e2.nopost = true;
                        e2.cbase = (rem >= 8 && rem < 20) ? (uint32_t)((rem - 8) & 3) : 0u;
                        uint32_t b = 0;
                        uint32_t r2 = slr_expand_group(tab, e2, vh, g, op, b);
                        loads++;
                        if (r2 != SLR_NONE32) r2 |= (uint32_t)nd << 8;
                        if (r2 < best2) { best2 = r2; bc2 = b; cnt2 = node_meta[nd] >> 10; }
                    }
                    if (best2 != SLR_NONE32 && (int)(best2 >> 8) * 21 + 21 <= base + 32) break;
                }
            }
        }
        if (best1 != SLR_NONE32) {
            const uint32_t c = slr_cnt_of(best1 & 15u);
            res.valid |= 1u; res.bc[0] = bc1;
            res.n_sub[0] = (uint8_t)(c & 3u); res.n_ins[0] = (uint8_t)((c >> 2) & 3u); res.n_del[0] = (uint8_t)((c >> 4) & 3u);
        }
        if (best2 != SLR_NONE32) {
            const uint32_t c = cnt2 + slr_cnt_of(best2 & 15u);
            res.valid |= 2u; res.bc[1] = bc2;
            res.n_sub[1] = (uint8_t)(c & 3u); res.n_ins[1] = (uint8_t)((c >> 2) & 3u); res.n_del[1] = (uint8_t)((c >> 4) & 3u);
        }
        out[qi] = res;
    }
    if (n_loads) *n_loads = loads;
    return 0;
}

// ---- UMI distance: same per-pair code as umi_dist.cu (bit planes of the row read, register-only Myers), pairs walked
// sequentially ----------------------------------------------------------------------------------------------------
#include "../../sicelore-2.1_b200/csrc/umi_core.cuh"
template <int L> static void sim_umi_job(const uint8_t *umis, int stride, long long n, int32_t *mat)
{
    for (long long i = 0; i < n; i++) {
        uint8_t rb[16] = {0};
        memcpy(rb, umis + i * stride, L + 2);
        uint32_t rw[4], pl[4];
        memcpy(rw, rb, 16);
        slr_umi_planes(rw, pl);
        const unsigned long long rp = slr_umi_planes_pack(pl);
        mat[i * n + i] = slr_umi_equality();
        for (long long v = i + 1; v < n; v++) {
            uint8_t cb[16] = {0};
            memcpy(cb, umis + v * stride, L + 2);
            uint32_t cw[4];
            memcpy(cw, cb, 16);
            const int32_t e = slr_umi_best9_planes<L>(rp, cw);
            mat[i * n + v] = e;
            mat[v * n + i] = slr_umi_transpose(e);
        }
    }
}
extern "C" int sim_umi_dist(const uint8_t *umis, int stride, int umi_len, const long long *job_offsets, long long n_jobs, int32_t *out,
                            const long long *out_offsets)
{
    for (long long j = 0; j < n_jobs; j++) {
        const long long j0 = job_offsets[j], n = job_offsets[j + 1] - j0;
        int32_t *mat = out + out_offsets[j];
        const uint8_t *u = umis + j0 * stride;
        switch (umi_len) {
#define C(LL) case LL: sim_umi_job<LL>(u, stride, n, mat); break;
            C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14)
#undef C
        default: return -1;
        }
    }
    return 0;
}
