// guided_core.cuh — per-lane logic of the Illumina-guided search kernel, __host__ __device__ so that
// tests/host_sim replays the same code on the CPU; the warp orchestration lives in guided_match.cu.
//
// Reference behaviour restated here (F! = Jar/NanoporeBC_UMI_finder-2.1.jar, T! = Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar):
//   BCUMIEDtesterBase.matchesSeqEditDistance / substitutions / insertions / deletions   F!…/TwoBit/ed/BCUMIEDtesterBase.class (BCUMIEDtesterBase.java:L82-L203)
//   NucTwoBitPerBaseEDtesterBase visited set + goNextEDlevel (bailout)                  F!…/TwoBit/ed/NucTwoBitPerBaseEDtesterBase.class (…java:L82-L144)
//   checkMatchWithTestSets: UMInucTwoBitPerBaseEDtester (…java:L52-L67), BCnucTwoBitPerBaseEDtester (…java:L72-L92)
//   reduction of the match list: IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI (…java:L52-L60), comparators (L106-L110, L143-L144)
//   NucleicAcidTwoBitPerBase replace/insert/delete                                      T!…/TwoBit/NucleicAcidTwoBitPerBase.class (…java:L228-L327)
//
// The engine enumerates the <= ed edit neighbourhood of one window depth first.  A deque node is (sequence, counters, previous /
// current position, level); popping it runs ONE position (<= 9 children: SUB A,G,C,T, INS A,G,C,T, DEL), every child is probed
// against the candidate sets and, below the last level, pushed.  The kernel keeps that order but runs all positions of a
// last-level node (whose children are never pushed) as one batch of 9·L children, 32 per warp step.
#pragma once
#include <stdint.h>
#include "../../include/sicelore_gpu.h"

#ifdef __CUDACC__
#define SLR_GHD static __host__ __device__ __forceinline__
#else
#define SLR_GHD static inline
#endif

constexpr int SLR_G_MAX_ED = 4;                 // IlluminaUMIanalyzer.java:L75 asserts ed < 5
constexpr int SLR_G_STACK = 48;                 // deque depth: (continuation + 9 children) per level
constexpr uint32_t SLR_G_EMPTY = 0xFFFFFFFFu;   // empty slot of a candidate table

// ---- candidate sets -------------------------------------------------------------------------------------------------------
// Every set (one per candidate group + the two global lists of the BC flavour) is a bucketised open-addressing table of 32-bit keys
// (L <= 16 => a clean 2-bit sequence fits 32 bits; values with garbage in bits 62-63 never match, see slr_g_child): a key hashes to
// a bucket of 8 slots = one 32-byte sector, read with two 128-bit loads and compared as a whole, so a lookup is one sector and the
// lanes of a warp do not diverge over probe chains (a full bucket without the key continues in the next one; at load <= 0.5 that is rare).
// meta: bits 0-4 log2(capacity in slots, >= 3), bit 8 "holds the all-T key 0xFFFFFFFF" (the empty marker), bit 9 "non-empty".
struct SlrGuidedSetsDev {
    const uint32_t *slots;          // every table starts on a 32-byte boundary
    const uint2 *groups;            // x = first slot, y = meta
    int n_groups;
    uint2 all_set, empty_set;       // BC flavour: allPassed10xBCs / outOfCellsBarcodes (meta 0 = absent)
    int all_ed, empty_ed;           // maxEDtoCheckBCAll10xBCs / maxEDtoCheckBCEmptyDrops
    int bc_flavour;
};

SLR_GHD uint32_t slr_g_hash(uint32_t key) { return key * 0x9E3779B1u; }
SLR_GHD uint32_t slr_g_bucket_of(uint32_t key, uint32_t lg_slots)          // lg_slots >= 3
{
    return lg_slots == 3u ? 0u : (slr_g_hash(key) >> (35u - lg_slots));     // top lg_slots - 3 bits
}

SLR_GHD bool slr_g_contains(const uint32_t *slots, uint2 set, uint32_t key)
{
    const uint32_t meta = set.y;
    if (!(meta & 0x200u)) return false;
    if (key == SLR_G_EMPTY) return (meta & 0x100u) != 0u;
    const uint32_t lg = meta & 31u, nb_mask = (1u << (lg - 3u)) - 1u;
    uint32_t bucket = slr_g_bucket_of(key, lg);
    const uint32_t *tab = slots + set.x;
    for (uint32_t i = 0; i <= nb_mask; i++) {
#if defined(__CUDA_ARCH__)
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(tab + 8u * bucket)), b = __ldg(reinterpret_cast<const uint4 *>(tab + 8u * bucket) + 1);
        const uint32_t v0 = a.x, v1 = a.y, v2 = a.z, v3 = a.w, v4 = b.x, v5 = b.y, v6 = b.z, v7 = b.w;
#else
        const uint32_t *q = tab + 8u * bucket;
        const uint32_t v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4], v5 = q[5], v6 = q[6], v7 = q[7];
#endif
        if (v0 == key || v1 == key || v2 == key || v3 == key || v4 == key || v5 == key || v6 == key || v7 == key) return true;
        if (v7 == SLR_G_EMPTY) return false;                  // slots of a bucket fill front to back: a free last slot ends the chain
        bucket = (bucket + 1u) & nb_mask;
    }
    return false;
}

// ---- visited set: IntHashSet on (int) seq of every node that has run a position (java:L105-L120), active for ed >= 2 ----
// per-warp table of 64-bit slots (stamp << 32 | key); a slot whose stamp is not the current window's is empty, so the table is
// never cleared between windows.
// Sizes: the set holds the DISTINCT sequences within ed-1 edits of the window that were expanded — measured with the oracle on
// random and homopolymer-rich 16-mers: <= 145 at ed 2, 4.1 k at ed 3, 17.7 k at ed 4 — tables of 2^9 / 2^15 / 2^17 slots.
SLR_GHD uint32_t slr_g_vis_log2(int ed) { return ed <= 2 ? 9u : (ed == 3 ? 15u : 17u); }

SLR_GHD bool slr_g_vis_contains(const unsigned long long *tab, uint32_t lg, uint32_t stamp, uint32_t key)
{
    const uint32_t mask = (1u << lg) - 1u;
    uint32_t slot = (slr_g_hash(key) >> (32u - lg)) & mask;
    for (uint32_t i = 0; i <= mask; i++) {                 // bounded: a full table cannot hang the warp
        const unsigned long long v = tab[slot];
        if ((uint32_t)(v >> 32) != stamp) return false;
        if ((uint32_t)v == key) return true;
        slot = (slot + 1u) & mask;
    }
    return false;
}

// false = the table is full (the read is then flagged SLR_G_TABLE_FULL; never observed)
SLR_GHD bool slr_g_vis_insert(unsigned long long *tab, uint32_t lg, uint32_t stamp, uint32_t key)
{
    const uint32_t mask = (1u << lg) - 1u;
    uint32_t slot = (slr_g_hash(key) >> (32u - lg)) & mask;
    for (uint32_t i = 0; i <= mask; i++) {
        const unsigned long long v = tab[slot];
        if ((uint32_t)(v >> 32) != stamp) { tab[slot] = ((unsigned long long)stamp << 32) | key; return true; }
        if ((uint32_t)v == key) return true;
        slot = (slot + 1u) & mask;
    }
    return false;
}

#if defined(__CUDACC__)
// concurrent insert by the lanes of one warp (batched nodes): the slot is claimed with a 64-bit CAS
static __device__ __forceinline__ bool slr_g_vis_insert_atomic(unsigned long long *tab, uint32_t lg, uint32_t stamp, uint32_t key)
{
    const uint32_t mask = (1u << lg) - 1u;
    uint32_t slot = (slr_g_hash(key) >> (32u - lg)) & mask;
    const unsigned long long want = ((unsigned long long)stamp << 32) | key;
    for (uint32_t i = 0; i <= mask;) {
        const unsigned long long v = *((volatile unsigned long long *)&tab[slot]);
        if (v == want) return true;
        if ((uint32_t)(v >> 32) != stamp) {
            if (atomicCAS(&tab[slot], v, want) == v) return true;
            continue;                                           // another lane took the slot: look at it again
        }
        slot = (slot + 1u) & mask;
        i++;
    }
    return false;
}
#endif

// ---- deque node (LongSeqMutated, LongSeqMutated.java:L44-L77) packed in 8 bytes ------------------------------------------------
// meta: bits 0-4 posTreatedInCurrentCycle + 1, 5-9 posTreatedInPreviousLevel + 1, 10-12 currentlevel, 13-15 nSubstitutions,
//       16-18 nInsertions, 19-21 nDeletions, 22 "dead" (bits 62-63 of the Java long are set), 23 findingErrorFlag GENE bit
struct SlrGNode { uint32_t seq, meta; };
SLR_GHD int  slr_g_pos_cur(uint32_t m)  { return (int)(m & 31u) - 1; }
SLR_GHD int  slr_g_pos_prev(uint32_t m) { return (int)((m >> 5) & 31u) - 1; }
SLR_GHD int  slr_g_level(uint32_t m)    { return (int)((m >> 10) & 7u); }
SLR_GHD int  slr_g_nsub(uint32_t m)     { return (int)((m >> 13) & 7u); }
SLR_GHD int  slr_g_nins(uint32_t m)     { return (int)((m >> 16) & 7u); }
SLR_GHD int  slr_g_ndel(uint32_t m)     { return (int)((m >> 19) & 7u); }
SLR_GHD bool slr_g_dead(uint32_t m)     { return (m >> 22) & 1u; }
SLR_GHD bool slr_g_inh(uint32_t m)      { return (m >> 23) & 1u; }
SLR_GHD uint32_t slr_g_root_meta()      { return (1u << 10); }                      // pos -1/-1, level 1, no errors (java:L82)
// goNextEDlevel (java:L137-L140): posPrev = parent's current position, posCur = -1, level + 1
SLR_GHD uint32_t slr_g_next_level_meta(uint32_t child_meta, int pos)
{
    return (child_meta & ~0x3FFu & ~(7u << 10)) | ((uint32_t)(pos + 1) << 5) | ((uint32_t)(slr_g_level(child_meta) + 1) << 10);
}

// ---- one child of node (seq, meta) at position p: j = 0-3 SUB A,G,C,T, 4-7 INS A,G,C,T, 8 DEL ---------------------------------
// low 32 bits of getLongHashReplaceByteDeg / InsertByteDeg / deleteByte for a sequence of L bases (digit i at bits 2(L-1-i));
// post2 holds the 2-bit value appended by a deletion for nDeletions = k at bits 2k (BYTE_TO_2BITLONG_ARRAY[0][getByteAt(k+1)]).
// valid  = the Java creates and probes this child, apart from the visited-set test;
// meta_out = the child's counters / flags (level and positions still the parent's: the Java copies the node);
// throws |= the deletion reads a post base outside ENCODE_MATRIX (postbad bit k = post base k is such a char):
//           BYTE_TO_2BITLONG_ARRAY[0][-1] -> ArrayIndexOutOfBoundsException, raised only when that base is used.
SLR_GHD uint32_t slr_g_child(uint32_t seq, uint32_t meta, int L, int p, int j, uint32_t post2, uint32_t postbad, int post_len, bool &valid,
                            uint32_t &meta_out, bool &throws)
{
    const int sh = 2 * (L - 1 - p);
    const uint32_t below = sh == 0 ? 0u : (0xFFFFFFFFu >> (32 - sh));      // digits p+1..L-1
    const uint32_t below2 = (below << 2) | 3u;                              // digits p..L-1
    const uint32_t b = (uint32_t)(j & 3);
    const int ndel = slr_g_ndel(meta);
    const uint32_t cbase = (post2 >> (2 * ndel)) & 3u;
    const uint32_t sub = (seq & ~(3u << sh)) | (b << sh);
    const uint32_t ins = (seq & ~below) | ((seq & below) >> 2) | ((b << sh) >> 2);
    const uint32_t del = (seq & ~below2) | ((seq << 2) & below2) | cbase;
    const bool is_sub = j < 4, is_ins = !is_sub && j < 8;
    valid = is_sub ? (((seq >> sh) & 3u) != b)            // s != cur.seq (java:L138)
          : is_ins ? (p < L - 1)                          // insertions only for posCur < L-1 (java:L115)
                   : (ndel + 1 <= post_len);              // deletions at every position, java:L187
    if (!is_sub && !is_ins && valid && ((postbad >> ndel) & 1u)) throws = true;
    // insert at p = L-2: `>>> 64` == `>>> 0` leaves the old last base in bits 62-63 (never matches; (int) value unaffected)
    const bool dead = slr_g_dead(meta) || (is_ins && p == L - 2 && (seq & 3u) != 0u);
    meta_out = (meta & ~(1u << 22)) + (is_sub ? (1u << 13) : is_ins ? (1u << 19) : (1u << 16)) | ((uint32_t)dead << 22);
    return is_sub ? sub : (is_ins ? ins : del);
}

// ---- candidate filter of a last-level node --------------------------------------------------------------------------------------
// Necessary condition for "c is a child of s" (one SUB, INS or DEL of the engine on an L-mer): c and s share a prefix, and the rest of c
// equals the rest of s either in place (SUB at p: prefix p, suffix L-1-p), shifted towards the end (INS after p: prefix p+1, then
// c[i] = s[i-1] for i >= p+2) or shifted towards the start (DEL at p: prefix p, c[i] = s[i+1] for p <= i <= L-2; the appended post base is
// not checked).  A last-level node none of whose candidates passes cannot produce a hit: its batch of 9·L children is skipped (the
// node still enters the visited set); otherwise only the positions that can hit are run.  Conservative by construction; tests
// enumerate every child of random nodes.
SLR_GHD int slr_g_clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
SLR_GHD int slr_g_ctz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;                // x != 0
#else
    return __builtin_ctz(x);
#endif
}
// pmin / pmax: the positions at which a child of s can equal c (SUB: the one differing digit; INS after p: prefix >= p+1 and shifted
// suffix from p+2; DEL at p: prefix >= p and shifted suffix from p) — the batch only has to run those positions.
SLR_GHD bool slr_g_may_be_child(uint32_t c, uint32_t s, int L, int &pmin, int &pmax)
{
    const uint32_t fm = L >= 16 ? 0xFFFFFFFFu : ((1u << (2 * L)) - 1u);
    c &= fm;
    const uint32_t x0 = (c ^ s) & fm;
    const int lead = x0 ? ((slr_g_clz32(x0) - (32 - 2 * L)) >> 1) : L;         // equal leading digits
    const int t0 = x0 ? (slr_g_ctz32(x0) >> 1) : L;                            // equal trailing digits
    const uint32_t x1 = (c ^ (s >> 2)) & fm;
    const int t1 = x1 ? (slr_g_ctz32(x1) >> 1) : L;                            // trailing digits with c[i] == s[i-1]
    const uint32_t x2 = ((c ^ (s << 2)) & fm) >> 2;
    const int t2 = (x2 ? (slr_g_ctz32(x2) >> 1) : L) + 1;                      // trailing digits with c[i] == s[i+1], the last one granted
    int lo = L, hi = -1;
    if (x0 && lead + t0 >= L - 1) { lo = lead; hi = lead; }                    // SUB at p = lead (x0 == 0: a substitution never recreates s)
    const int ilo = L - 2 - t1 > 0 ? L - 2 - t1 : 0, ihi = lead - 1 < L - 2 ? lead - 1 : L - 2;
    if (ilo <= ihi) { lo = ilo < lo ? ilo : lo; hi = ihi > hi ? ihi : hi; }    // INS after p in [L-2-t1, lead-1]
    const int dlo = L - t2 > 0 ? L - t2 : 0, dhi = lead < L - 1 ? lead : L - 1;
    if (dlo <= dhi) { lo = dlo < lo ? dlo : lo; hi = dhi > hi ? dhi : hi; }    // DEL at p in [L-t2, lead]
    pmin = lo; pmax = hi;
    return lo <= hi;
}
// slot `i` of a small group's table as a filter candidate: valid unless the slot is the empty marker of a group without the all-T key
SLR_GHD bool slr_g_filter_slot(const uint32_t *slots, uint2 set, uint32_t i, uint32_t &cand)
{
    const uint32_t meta = set.y;
    cand = 0u;
    if (!(meta & 0x200u) || i >= (1u << (meta & 31u))) return false;
#if defined(__CUDA_ARCH__)
    cand = __ldg(slots + set.x + i);
#else
    cand = slots[set.x + i];
#endif
    return cand != SLR_G_EMPTY || (meta & 0x100u) != 0u;
}
SLR_GHD bool slr_g_filter_usable(uint2 set) { return !(set.y & 0x200u) || (set.y & 31u) <= 6u; }    // empty group or <= 64 slots

// ---- "far node" test of a node one level above the last ------------------------------------------------------------------------
// One engine operation x -> x' is one Levenshtein edit between prefixes: SUB Lev(x, x') = 1, INS Lev(x[:L-1], x') = 1 (the last base is
// pushed out), DEL Lev(x, x'[:L-1]) = 1 (an unchecked base is pulled in).  Cutting one more base off either side of such a pair keeps the
// distance <= 1 for a suitable cut of the other side, so a candidate c reachable from s with <= 2 operations satisfies
//     Lev(s[:i], c[:j]) <= 2   for some i, j in L-2..L
// (necessary, not sufficient: the engine's visited set, position skip and dead values only remove paths).  Myers / Hyyro bit-parallel
// global edit distance, pattern = candidate (its four match masks are per-read constants), text = node: the last-row score and the two top
// vertical deltas after the columns L-2..L give the nine values.  A node that is far from every candidate cannot produce a hit in its
// whole subtree.
struct SlrGPeq { uint32_t eq[4]; };
SLR_GHD SlrGPeq slr_g_peq(uint32_t c, int L)
{
    SlrGPeq P;
    P.eq[0] = P.eq[1] = P.eq[2] = P.eq[3] = 0u;
    for (int i = 0; i < L; i++) {
        const uint32_t d = (c >> (2 * (L - 1 - i))) & 3u;        // pattern position i (0 = first base) -> bit i
        P.eq[0] |= (uint32_t)(d == 0u) << i; P.eq[1] |= (uint32_t)(d == 1u) << i;
        P.eq[2] |= (uint32_t)(d == 2u) << i; P.eq[3] |= (uint32_t)(d == 3u) << i;
    }
    return P;
}
SLR_GHD int slr_g_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
SLR_GHD bool slr_g_within2(const SlrGPeq &P, uint32_t s, int L)
{
    const uint32_t ones = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    uint32_t Pv = ones, Mv = 0u;
    int score = L, best = 99;                                     // score = D[L][j] = Lev(c, s[:j])
    for (int j = 0; j < L; j++) {
        const uint32_t d = (s >> (2 * (L - 1 - j))) & 3u;
        const uint32_t Eq = d == 0u ? P.eq[0] : d == 1u ? P.eq[1] : d == 2u ? P.eq[2] : P.eq[3];
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = ((((Eq & Pv) + Pv) ^ Pv) | Eq) & ones;
        uint32_t Ph = (Mv | ~(Xh | Pv)) & ones;
        uint32_t Mh = Pv & Xh;
        score += (int)((Ph >> (L - 1)) & 1u) - (int)((Mh >> (L - 1)) & 1u);
        Ph = ((Ph << 1) | 1u) & ones;                              // D[0][j] = j
        Mh = (Mh << 1) & ones;
        Pv = (Mh | ~(Xv | Ph)) & ones;
        Mv = Ph & Xv;
        if (j >= L - 3) {                                          // columns L-2..L: D[L][.], D[L-1][.], D[L-2][.]
            const int d1 = score - (int)((Pv >> (L - 1)) & 1u) + (int)((Mv >> (L - 1)) & 1u);
            const int d2 = d1 - (int)((Pv >> (L - 2)) & 1u) + (int)((Mv >> (L - 2)) & 1u);
            const int m = score < d1 ? (score < d2 ? score : d2) : (d1 < d2 ? d1 : d2);
            if (m < best) best = m;
        }
    }
    return best <= 2;
}

// ---- checkMatchWithTestSets: 0 = no hit, else SLR_G_W_* bits of the list entry; inh_out = GENE bit now on the node ------
SLR_GHD uint32_t slr_g_probe(const SlrGuidedSetsDev &S, uint2 group, uint32_t s, uint32_t meta, int level, bool &inh_out)
{
    inh_out = slr_g_inh(meta);
    if (slr_g_dead(meta)) return 0u;
    if (slr_g_contains(S.slots, group, s)) {
        if (!S.bc_flavour) return 0x80u;                                    // UMI flavour: a hit without flag bits
        inh_out = true;                                                     // BCnuc…java:L76: on the node itself
        return SLR_G_W_GENE;
    }
    if (!S.bc_flavour) return 0u;
    const uint32_t g = inh_out ? SLR_G_W_GENE : 0u;
    if (level <= S.all_ed && slr_g_contains(S.slots, S.all_set, s)) return g | SLR_G_W_ALL;        // java:L79-L83
    if (level <= S.empty_ed && slr_g_contains(S.slots, S.empty_set, s)) return g | SLR_G_W_EMPTY;  // java:L85-L89
    return 0u;
}

// ---- running "sorted().distinct()" of the match list: the first two entries -------------------------------------------------
// key = (getNErrors, [scoreWhereFound], |offset|); hits arrive in list order, so an equal key never displaces an earlier hit
// (Stream.sorted is stable) and distinct() keeps the first entry of each sequence.
struct SlrGTop2 {
    uint32_t seq[2], key[2], info[2];
    int n, n_raw, min_err_gene;
};
SLR_GHD void slr_g_top2_init(SlrGTop2 &T) { T.n = 0; T.n_raw = 0; T.min_err_gene = 0x7FFFFFFF; T.seq[0] = T.seq[1] = 0; T.key[0] = T.key[1] = 0xFFFFFFFFu; T.info[0] = T.info[1] = 0; }
SLR_GHD uint32_t slr_g_score(uint32_t where)                                  // BarcodeFindingFlag.java:L119-L131
{
    return (where & SLR_G_W_GENE) ? 3u : (where & SLR_G_W_ALL) ? 2u : (where & SLR_G_W_EMPTY) ? 1u : 0u;
}
SLR_GHD void slr_g_top2_add(SlrGTop2 &T, int bc_flavour, uint32_t seq, uint32_t child_meta, int offset, uint32_t where)
{
    where &= 7u;
    const int nerr = slr_g_nsub(child_meta) + slr_g_nins(child_meta) + slr_g_ndel(child_meta);
    const uint32_t key = ((uint32_t)nerr << 8) | ((bc_flavour ? slr_g_score(where) : 0u) << 4) | (uint32_t)(offset < 0 ? -offset : offset);
    const uint32_t info = (uint32_t)slr_g_nsub(child_meta) | ((uint32_t)slr_g_nins(child_meta) << 4) | ((uint32_t)slr_g_ndel(child_meta) << 8) |
                          ((uint32_t)(offset + 8) << 12) | (where << 16);
    T.n_raw++;
    if ((where & SLR_G_W_GENE) && nerr < T.min_err_gene) T.min_err_gene = nerr;
    if (T.n == 0) { T.seq[0] = seq; T.key[0] = key; T.info[0] = info; T.n = 1; return; }
    if (seq == T.seq[0]) {
        if (key < T.key[0]) { T.key[0] = key; T.info[0] = info; }
        return;
    }
    if (key < T.key[0]) {
        T.seq[1] = T.seq[0]; T.key[1] = T.key[0]; T.info[1] = T.info[0];
        T.seq[0] = seq; T.key[0] = key; T.info[0] = info; T.n = 2;
        return;
    }
    if (T.n == 1 || key < T.key[1]) { T.seq[1] = seq; T.key[1] = key; T.info[1] = info; T.n = 2; }
}
SLR_GHD void slr_g_top2_store(const SlrGTop2 &T, uint32_t flags, slr_guided_result &r)
{
    for (int i = 0; i < 2; i++) {
        const bool on = i < T.n && !flags;
        const uint32_t f = on ? T.info[i] : 0u;
        r.seq[i] = on ? (uint64_t)T.seq[i] : 0ull;
        r.n_sub[i] = (int8_t)(f & 15u); r.n_ins[i] = (int8_t)((f >> 4) & 15u); r.n_del[i] = (int8_t)((f >> 8) & 15u);
        r.offset[i] = on ? (int8_t)((int)((f >> 12) & 15u) - 8) : (int8_t)0;
        r.where[i] = (uint8_t)((f >> 16) & 7u);
    }
    r.n_distinct = flags ? (uint8_t)0 : (uint8_t)T.n;
    r.flags = (uint8_t)flags;
    r.n_raw = flags ? 0 : T.n_raw;
    r.min_err_gene = flags ? 0x7FFFFFFF : T.min_err_gene;
    r.pad = 0;
}

// ---- characters ----------------------------------------------------------------------------------------------------------------
// NucleicAcidByteCodeBase.ENCODE_MATRIX (T!…NucleicAcidByteCodeBase.java:L45-L78): 4-bit code, 0xFF = not in the matrix
SLR_GHD uint32_t slr_g_code4(uint32_t c)
{
    if (c == '-') return 0u;
    switch (c | 0x20u) {
    case 'a': return 1u; case 'g': return 2u; case 'c': return 4u; case 't': return 8u; case 'n': return 15u;
    case 'h': return 13u; case 'r': return 3u; case 'y': return 12u; case 'm': return 5u; case 'k': return 10u;
    case 's': return 6u; case 'w': return 9u; case 'b': return 14u; case 'v': return 7u; case 'd': return 11u;
    default: return 0xFFu;
    }
}
// FOURBIT_TO_TWOBIT_MATRIX / BYTE_TO_2BITLONG_ARRAY[0]: G->1, C->2, T->3, every other code (A, IUPAC, N in a post sequence) -> 0
SLR_GHD uint32_t slr_g_two_of_code4(uint32_t c4) { return c4 == 2u ? 1u : c4 == 4u ? 2u : c4 == 8u ? 3u : 0u; }
SLR_GHD int slr_g_offset_of(int k) { return (k == 0) ? 0 : ((k & 1) ? -((k + 1) / 2) : (k / 2)); }   // 0,-1,1,-2,2 (java:L89-L91)
