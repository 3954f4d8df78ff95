// guided_sim.cu — TEST INFRASTRUCTURE: replays the warp orchestration of guided_match.cu sequentially on the CPU with the very same
// per-lane code (guided_core.cuh) and the same table builder (guided_build.h), so that the algorithm (leaf batches, self rule,
// stamped visited table, bailout ranks, running top-2) is checked against the oracle in the GPU-less container.
#include <cstring>
#include <vector>
#include "../../sicelore-2.1_b200/csrc/guided_build.h"

namespace {

int sim_no_filter = 0;             // sim_guided_set_filter(0): run every last-level batch (checks that skipping is exact)
long long sim_filter_skips = 0, sim_far_nodes = 0;

void sim_record(SlrGTop2 &T, int bc_flavour, uint32_t seq, uint32_t cmeta, int level, int offset, uint32_t where, slr_guided_hit *raw, int raw_cap)
{
    if (raw && T.n_raw < raw_cap) {
        slr_guided_hit h;
        h.seq = seq; h.n_sub = (int8_t)slr_g_nsub(cmeta); h.n_ins = (int8_t)slr_g_nins(cmeta); h.n_del = (int8_t)slr_g_ndel(cmeta);
        h.offset = (int8_t)offset; h.where = (uint8_t)(where & 7u); h.level = (uint8_t)level; h.pad = 0;
        raw[T.n_raw] = h;
    }
    slr_g_top2_add(T, bc_flavour, seq, cmeta, offset, where);
}

void sim_query(const SlrGuidedSetsDev &S, int L, int plusminus, int post_len, int bailout, const uint8_t *slice, int slice_len, int anc, int gid,
               int ed, int max_ed, slr_guided_result *out, slr_guided_hit *raw, int raw_cap, unsigned long long *vis, uint32_t &stamp)
{
    uint8_t codes[32];
    for (int lane = 0; lane < 32; lane++) codes[lane] = (uint8_t)slr_g_code4(lane < slice_len ? slice[lane] : 0u);
    const uint2 group = (gid >= 0 && gid < S.n_groups) ? S.groups[gid] : make_uint2(0u, 0u);
    const uint32_t vlg = slr_g_vis_log2(ed);
    const bool use_vis = ed >= 2;
    const int nchild = 9 * L;
    uint32_t cand[64];
    bool cok[64];
    for (uint32_t i = 0; i < 64; i++) cok[i] = slr_g_filter_slot(S.slots, group, i, cand[i]);
    const bool filt_leaf = slr_g_filter_usable(group) &&
                           !(S.bc_flavour && (((S.all_set.y & 0x200u) && ed <= S.all_ed) || ((S.empty_set.y & 0x200u) && ed <= S.empty_ed)));
    const bool filt_inner = ed >= 2 && slr_g_filter_usable(group) &&
                            !(S.bc_flavour && (((S.all_set.y & 0x200u) && ed - 1 <= S.all_ed) || ((S.empty_set.y & 0x200u) && ed - 1 <= S.empty_ed)));
    SlrGPeq peq[64];
    if (filt_inner) for (int i = 0; i < 64; i++) peq[i] = slr_g_peq(cand[i], L);
    SlrGTop2 T;
    slr_g_top2_init(T);
    uint32_t flags = (ed < 0 || ed > max_ed) ? SLR_G_EXCEPTION : 0u;
    SlrGNode stack[SLR_G_STACK];
    for (int k = 0; k <= 2 * plusminus && !flags; k++) {
        const int o = slr_g_offset_of(k), ws = anc + o;
        if (ws < 0 || ws + L + post_len > slice_len) { flags = SLR_G_EXCEPTION; break; }
        uint32_t w = 0, post2 = 0, postbad = 0;
        bool bad = false, throws = false;
        for (int lane = 0; lane < L + post_len; lane++) {
            const uint32_t c4 = codes[ws + lane];
            if (lane < L && c4 >= 15u) bad = true;
            if (lane >= L && c4 == 0xFFu) postbad |= 1u << (lane - L);
            const uint32_t two = slr_g_two_of_code4(c4);
            if (lane < L) w |= two << (2 * (L - 1 - lane));
            else if (lane - L < 16) post2 |= two << (2 * (lane - L));
        }
        if (bad) { flags = SLR_G_EXCEPTION; break; }
        stamp++;
        int nlist = 0;
        uint32_t root_meta = slr_g_root_meta();
        {
            bool inh;
            const uint32_t where = slr_g_probe(S, group, w, root_meta, 1, inh);
            if (where) { sim_record(T, S.bc_flavour, w, root_meta, 1, o, where, raw, raw_cap); nlist++; }
            if (inh) root_meta |= 1u << 23;
        }
        if (ed == 0) continue;
        int sp = 0;
        stack[sp].seq = w; stack[sp].meta = root_meta; sp++;
        while (sp > 0) {
            const SlrGNode node = stack[--sp];
            const int level = slr_g_level(node.meta), pos_prev = slr_g_pos_prev(node.meta);
            if (level == ed) {
                const int p0 = pos_prev == 0 ? 1 : 0;
                int c_lo = 0, c_hi = nchild - 1;
                bool run = !slr_g_dead(node.meta);
                if (run && filt_leaf && postbad == 0u && !sim_no_filter) {
                    int pmin = L, pmax = -1;
                    for (int i = 0; i < 64; i++) {
                        int a = L, b = -1;
                        if (cok[i]) slr_g_may_be_child(cand[i], node.seq, L, a, b);
                        if (a < pmin) pmin = a;
                        if (b > pmax) pmax = b;
                    }
                    run = pmin <= pmax;
                    c_lo = 9 * pmin; c_hi = 9 * pmax + 8;
                    sim_filter_skips += !run;
                }
                for (int c0 = c_lo; run && c0 <= c_hi; c0 += 32) {
                    uint32_t s_[32], cm_[32], wh_[32];
                    for (int lane = 0; lane < 32; lane++) {          // all lanes of a step see the same visited table
                        const int c = c0 + lane, p = c / 9, j = c - 9 * p;
                        bool valid = false, inh;
                        uint32_t cmeta = 0, s = 0, where = 0;
                        if (c <= c_hi && p != pos_prev) {
                            s = slr_g_child(node.seq, node.meta, L, p, j, post2, postbad, post_len, valid, cmeta, throws);
                            if (valid && use_vis && ((s == node.seq && p > p0) || slr_g_vis_contains(vis, vlg, stamp, s))) valid = false;
                            if (valid) where = slr_g_probe(S, group, s, cmeta, level, inh);
                        }
                        s_[lane] = s; cm_[lane] = cmeta; wh_[lane] = where;
                    }
                    for (int lane = 0; lane < 32; lane++)
                        if (wh_[lane]) { sim_record(T, S.bc_flavour, s_[lane], cm_[lane], level, o, wh_[lane], raw, raw_cap); nlist++; }
                }
                if (use_vis && !slr_g_vis_insert(vis, vlg, stamp, node.seq)) flags |= SLR_G_TABLE_FULL;
                continue;
            }
            if (level == ed - 1 && slr_g_pos_cur(node.meta) < 0 && filt_inner && postbad == 0u && !sim_no_filter) {
                bool far = slr_g_dead(node.meta);
                if (!far) {
                    far = true;
                    for (int i = 0; i < 64; i++) if (cok[i] && slr_g_within2(peq[i], node.seq, L)) far = false;
                }
                if (far) {
                    sim_far_nodes++;
                    if (bailout < 0 || level < bailout || nlist == 0)
                        for (int c = 0; c < nchild; c++) {
                            const int p = c / 9, j = c - 9 * p;
                            if (p == pos_prev) continue;
                            bool valid = false, thr = false;
                            uint32_t cmeta;
                            const uint32_t sc = slr_g_child(node.seq, node.meta, L, p, j, post2, 0u, post_len, valid, cmeta, thr);
                            if (valid && !slr_g_vis_insert(vis, vlg, stamp, sc)) flags |= SLR_G_TABLE_FULL;
                        }
                    if (!slr_g_vis_insert(vis, vlg, stamp, node.seq)) flags |= SLR_G_TABLE_FULL;
                    continue;
                }
            }
            const int pos = slr_g_pos_cur(node.meta) + 1;
            const uint32_t meta = (node.meta & ~31u) | (uint32_t)(pos + 1);
            if (pos < L - 1) { stack[sp].seq = node.seq; stack[sp].meta = meta; sp++; }
            if (pos_prev == pos) continue;
            uint32_t s_[9], cm_[9], wh_[9];
            bool valid_[9], inh_[9];
            for (int lane = 0; lane < 9; lane++) {
                bool valid = false, inh = false;
                uint32_t cmeta = 0, where = 0;
                uint32_t s = slr_g_child(node.seq, meta, L, pos, lane, post2, postbad, post_len, valid, cmeta, throws);
                if (valid && use_vis && slr_g_vis_contains(vis, vlg, stamp, s)) valid = false;
                if (valid) where = slr_g_probe(S, group, s, cmeta, level, inh);
                s_[lane] = s; cm_[lane] = cmeta; wh_[lane] = where; valid_[lane] = valid; inh_[lane] = inh;
            }
            int hits_le = 0;
            bool push_[9];
            for (int lane = 0; lane < 9; lane++) {
                if (wh_[lane]) hits_le++;
                push_[lane] = valid_[lane] && (bailout < 0 || level < bailout || nlist + hits_le == 0);
            }
            for (int lane = 0; lane < 9; lane++)
                if (wh_[lane]) { sim_record(T, S.bc_flavour, s_[lane], cm_[lane], level, o, wh_[lane], raw, raw_cap); nlist++; }
            for (int lane = 0; lane < 9; lane++)
                if (push_[lane]) {
                    stack[sp].seq = s_[lane];
                    stack[sp].meta = slr_g_next_level_meta((cm_[lane] & ~(1u << 23)) | ((uint32_t)inh_[lane] << 23), pos);
                    sp++;
                }
            if (use_vis && !slr_g_vis_insert(vis, vlg, stamp, node.seq)) flags |= SLR_G_TABLE_FULL;
        }
        if (throws) flags = SLR_G_EXCEPTION;
    }
    slr_g_top2_store(T, flags, *out);
}

}  // namespace

extern "C" void sim_guided_batch(const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups, const uint64_t *all_keys,
                                 int64_t n_all, int all_ed, const uint64_t *empty_keys, int64_t n_empty, int empty_ed, int bc_flavour, int L,
                                 int plusminus, int bailout, int post_len, const uint8_t *slices, int stride, int slice_len,
                                 const int32_t *anchor, const int32_t *group_id, const int32_t *ed, int64_t n, slr_guided_result *out,
                                 slr_guided_hit *raw_out, int raw_cap)
{
    SlrGuidedSetsHost H;
    slr_guided_build(group_keys, group_offsets, n_groups, all_keys, n_all, empty_keys, n_empty, L, H);
    SlrGuidedSetsDev S;
    S.slots = H.slots.data(); S.groups = H.groups.data(); S.n_groups = (int)n_groups;
    S.all_set = H.all_set; S.empty_set = H.empty_set; S.all_ed = all_ed; S.empty_ed = empty_ed; S.bc_flavour = bc_flavour ? 1 : 0;
    int max_ed = 0;
    for (int64_t i = 0; i < n; i++) if (ed[i] > max_ed) max_ed = ed[i];
    if (max_ed > SLR_G_MAX_ED) max_ed = SLR_G_MAX_ED;
    std::vector<unsigned long long> vis((size_t)1 << slr_g_vis_log2(max_ed), 0ull);   // one "warp": the table is shared by all queries, like on the GPU
    uint32_t stamp = 0;
    if (raw_out) memset(raw_out, 0, (size_t)n * raw_cap * sizeof(slr_guided_hit));
    for (int64_t i = 0; i < n; i++)
        sim_query(S, L, plusminus, post_len, bailout, slices + (size_t)i * stride, slice_len, anchor[i], group_id[i], ed[i], max_ed, &out[i],
                  raw_out ? raw_out + (size_t)i * raw_cap : nullptr, raw_cap, vis.data(), stamp);
}

extern "C" void sim_guided_set_filter(int on) { sim_no_filter = !on; }
extern "C" long long sim_guided_far_nodes(int reset) { const long long r = sim_far_nodes; if (reset) sim_far_nodes = 0; return r; }
extern "C" long long sim_guided_filter_skips(int reset) { const long long r = sim_filter_skips; if (reset) sim_filter_skips = 0; return r; }

// every child the engine can create from `n` random nodes must pass slr_g_may_be_child: returns the number of violations
extern "C" long long sim_guided_filter_violations(int L, int post_len, long long n, unsigned long long seed)
{
    unsigned long long x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 16); };
    const uint32_t fm = L >= 16 ? 0xFFFFFFFFu : ((1u << (2 * L)) - 1u);
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        uint32_t s = rnd() & fm;
        if (i % 3 == 0) s &= rnd() & rnd();                    // homopolymer-rich
        const uint32_t post2 = rnd();
        const uint32_t meta = slr_g_root_meta() | ((rnd() % 3u) << 19);     // nDeletions 0..2 picks the appended post base
        for (int p = 0; p < L; p++)
            for (int j = 0; j < 9; j++) {
                bool valid, throws = false;
                uint32_t cm;
                const uint32_t c = slr_g_child(s, meta, L, p, j, post2, 0u, post_len, valid, cm, throws);
                int a = L, b = -1;
                if (valid && !(slr_g_may_be_child(c, s, L, a, b) && a <= p && p <= b)) bad++;      // passes, and p lies in the reported range
            }
    }
    return bad;
}
