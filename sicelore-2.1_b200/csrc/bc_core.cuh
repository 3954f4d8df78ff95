// bc_core.cuh — per-lane logic of the barcode-assignment kernel, __host__ __device__ so that
// tests/host_sim replays exactly the same code on the CPU (the warp orchestration lives in bc_assign.cu).
//
// Reference behaviour restated here (F! = Jar/NanoporeBC_UMI_finder-2.1.jar, T! = Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar):
//   BarcodeMatchTester.doJob / substitutions / insertions / deletions   F!…/analyzers/BarcodeMatchTester.class (BarcodeMatchTester.java:L198-L357)
//   NucTwoBitPerBaseEDtesterBase visited set                            F!…/TwoBit/ed/NucTwoBitPerBaseEDtesterBase.class (…java:L82-L120)
//   NucleicAcidTwoBitPerBase replace/insert/delete                      T!…/TwoBit/NucleicAcidTwoBitPerBase.class (…java:L228-L327)
//   Parser.assignBarcode merge + decision                               F!…/analyzers/Parser.class (Parser.java:L240-L311)
#pragma once
#include "slr_table.cuh"
#include "../../include/sicelore_gpu.h"

constexpr int SLR_MAX_OFFSETS = 9;             // plusminus <= 4
constexpr uint32_t SLR_NONE32 = 0xFFFFFFFFu;
constexpr unsigned long long SLR_VH_EMPTY = 0xFFFFFFFFFFFFFFFFull;
#ifndef SLR_VH_BITS
#define SLR_VH_BITS 9                          // 512 slots for at most 144 values: at a load of 0.28 the warp-wide insert rarely probes twice
#endif
constexpr int SLR_VH_SIZE = 1 << SLR_VH_BITS;

// matches of one read: one slot per (offset index, ED level) — Matches is a HashSet whose equals() is
// (readSeq, ED, offset) (BarcodeMatchTester.java:L433-L436), i.e. first hit per ED level per offset wins.
struct SlrMatchStore {
    uint32_t m_w[SLR_MAX_OFFSETS];             // window = OneMatch.readSeq
    uint32_t m_bc[SLR_MAX_OFFSETS][3];         // OneMatch.matchingBC
    uint8_t m_cnt[SLR_MAX_OFFSETS][3];         // nSub | nIns << 2 | nDel << 4
    uint8_t m_valid[SLR_MAX_OFFSETS];          // bit ED set = that level has a match
};

SLR_HD int slr_offset_of(int k) { return (k == 0) ? 0 : ((k & 1) ? -((k + 1) / 2) : (k / 2)); }   // 0,-1,1,-2,2 (Parser.java:L198-L200)

// ---- visited-hash lookup (value -> earliest processing time); insertion is done by the orchestration ----
SLR_HD uint32_t slr_vh_slot(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - SLR_VH_BITS); }
SLR_HD uint32_t slr_vh_tmin(const unsigned long long *tab, uint32_t v)
{
    uint32_t slot = slr_vh_slot(v);
    while (true) {
        const unsigned long long cur = tab[slot];
        if (cur == SLR_VH_EMPTY) return SLR_NONE32;
        if ((uint32_t)(cur >> 32) == v) return (uint32_t)cur;
        slot = (slot + 1) & (SLR_VH_SIZE - 1);
    }
}

// ---- level-1 mutant of window w: root position p, creation index j (0-3 SUB A,G,C,T; 4-7 INS; 8 DEL) ----
// low 32 bits of getLongHashReplaceByteDeg / InsertByteDeg / deleteByte (java:L228-L234, L300-L310, L321-L327).
// dead = the Java value carries garbage in bits 62-63 (insert at p = L-2 shifts by 64 == 0): never matches,
// but its (int) value still enters the visited set.  Straight-line (lanes of one warp mix all three kinds).
SLR_HD uint32_t slr_gen_mutant(uint32_t w, int p, int j, uint32_t cbase, bool &valid, bool &dead)
{
    const int sh = 2 * (15 - p);
    const uint32_t below = slr_lowmask(sh);                    // digits p+1..15
    const uint32_t below2 = (below << 2) | 3u;                 // digits p..15
    const uint32_t sub = (w & ~(3u << sh)) | ((uint32_t)(j & 3) << sh);
    const uint32_t ins = (w & ~below) | ((w & below) >> 2) | (((uint32_t)(j & 3) << sh) >> 2);
    const uint32_t del = (w & ~below2) | ((w << 2) & below2) | cbase;
    const bool is_sub = j < 4;
    valid = is_sub ? (((w >> sh) & 3u) != (uint32_t)j)          // s != cur.seq (L260)
                   : (p < 15);                                  // indels only for posCur < L-1 (L234)
    dead = (!is_sub) && (j < 8) && (p == 14) && ((w & 3u) != 0u);
    return is_sub ? sub : (j < 8 ? ins : del);
}

// postSeq == null variant (BarcodeDatasetColissionTester): 12 creations per position, j = 8 + b is the deletion that
// appends base b (deleteByte with code 0, then | 1, 2, 3: BarcodeMatchTester.java:L330-L340)
SLR_HD uint32_t slr_gen_mutant12(uint32_t w, int p, int j, bool &valid, bool &dead)
{
    return slr_gen_mutant(w, p, j < 8 ? j : 8, j < 8 ? 0u : (uint32_t)(j - 8), valid, dead);
}

// Context of one node expansion (= all positions of one LongSeqMutated popped from the deque)
struct SlrExpand {
    uint32_t cs;        // node sequence
    uint32_t w;         // unmutated window
    int pskip;          // posTreatedInPreviousLevel (must not be mutated again, L227), -1 for the root
    uint32_t cbase;     // base appended by a deletion: post[nDel+1] (L329)
    uint32_t tproc;     // processing time of this node (level 2 only)
    int level;          // 1 = root expansion (hits are ED 1), 2 = level-1 node expansion (hits are ED 2)
    bool use_visited;   // ed >= 2 (NucTwoBitPerBaseEDtesterBase.java:L82-L95)
    bool nopost;        // postSeq == null (pass-1 collision tester): a deletion appends all four bases (L336-L340), creation
                        // index 8 + base; cbase then names the appended base of THIS probe (digit groups 0-2)
};

// Would the reference have skipped mutant s, created at position q of this node, as "already tested"?
// visited (java:L105-L120) holds the (int) seq of every node that finished at least one position:
//   level 1: {w, once the root did position 0} U {level-1 mutants of earlier root positions}
//   level 2: {w} U {level-1 nodes processed before this node} U {this node, after its first position}
// Monotone in q: visited at q implies visited at every later position of the same node.
// Level 1 needs no table: only the FIRST hit of the root expansion is kept and the earliest creation of a value is
// never "already tested", so a later duplicate can only matter when every earlier creation was a dead one (the
// INS at p = L-2 whose Java value carries garbage in bits 62-63: it never matches, but its (int) value
// (w & ~3) | b is marked).  The only later creations of those values are the substitutions at the last position.
SLR_HD bool slr_is_visited(const SlrExpand &e, const unsigned long long *vh, uint32_t s, int q)
{
    if (e.level == 1) {
        if (!e.use_visited) return false;
        if (q >= 1 && s == e.w) return true;
        return q == 15 && (e.w & 3u) != 0u && ((s ^ e.w) & ~3u) == 0u;
    }
    if (s == e.w) return true;
    const uint32_t t = slr_vh_tmin(vh, s);
    if (t == SLR_NONE32) return false;
    if (t < e.tproc) return true;
    const int firstpos = (e.pskip == 0) ? 1 : 0;
    return t == e.tproc && q > firstpos;
}

// Test slot pattern P of table g against the op-mutants (0 SUB, 1 INS, 2 DEL) of node e that fall into digit
// group g.  Returns the smallest traversal rank q*16 + idx (idx: 0-3 SUB base, 4-7 INS base, 8 DEL) of a
// generating, non-visited mutant, or SLR_NONE32.  rest = the probe's rest (s = slr_key_join(rest, P, g) is the
// full candidate barcode).
// INS / DEL: "a 4-digit string is a 3-digit string plus one digit": with lpre = common leading digits and
// lsuf = common trailing digits the generating positions are exactly the interval [3 - lsuf, lpre].
// The first half is a straight-line filter (a slot with the right tag is a list barcode that agrees with the node
// outside the digit group, but only ~5 % of those are one edit away); the rest runs for real candidates only.
SLR_HD uint32_t slr_check_pattern(const SlrExpand &e, const unsigned long long *vh, int g, int op, uint32_t P,
                                                      uint32_t rest, uint32_t &s_out)
{
    const uint32_t csg = (e.cs >> (24 - 8 * g)) & 0xFFu;
    const uint32_t x = P ^ csg;
    const uint32_t d = (x | (x >> 1)) & 0x55u;
    const bool sub_ok = d != 0u && (d & (d - 1u)) == 0u;        // exactly one digit differs
    const uint32_t long4 = (op == 1) ? P : csg, short3 = (op == 1) ? (csg >> 2) : (P >> 2);
    const uint32_t xh = (long4 >> 2) ^ short3, xl = (long4 & 0x3Fu) ^ short3;
    const int hi = (slr_clz(xh) - 26) >> 1;                    // common leading digits, 0..3
    int xx = 3 - ((slr_ffs(xl | 0x40u) - 1) >> 1);             // 3 - common trailing digits
    const uint32_t n0 = (g < 3) ? ((e.cs >> (22 - 8 * g)) & 3u) : e.cbase;   // digit that follows the group after a deletion
    const bool indel_ok = xx <= hi && (op == 1 || (P & 3u) == n0 || (e.nopost && g == 3));
    if (!(op == 0 ? sub_ok : indel_ok)) return SLR_NONE32;

    int q;
    uint32_t idx;
    if (op == 0) {                                             // substitutions (L257-L273)
        const int il = 3 - ((slr_ffs(d) - 1) >> 1);
        q = 4 * g + il;
        if (q == e.pskip) return SLR_NONE32;
        idx = (P >> (2 * (3 - il))) & 3u;
    } else {                                                   // insertions (L284-L300): new digit at j = q+1; deletions (L313-L357)
        q = -2;
        for (; xx <= hi; xx++) {                               // almost always a single candidate
            const int qq = (op == 1) ? 4 * g + xx - 1 : 4 * g + xx;
            if (qq < 0 || qq > 14 || qq == e.pskip) continue;
            if (op == 1 && qq == 14 && (e.cs & 3u) != 0u) continue;   // the `>>> 64` value: garbage in bits 62-63
            q = qq;
            break;
        }
        if (q < 0) return SLR_NONE32;
        idx = (op == 1) ? 4u + ((P >> (2 * (3 - xx))) & 3u) : (e.nopost ? 8u + (g == 3 ? (P & 3u) : e.cbase) : 8u);
    }
    const uint32_t s = slr_key_join(rest, P, g);
    if (slr_is_visited(e, vh, s, q)) return SLR_NONE32;
    s_out = s;
    return (uint32_t)(q * 16) + idx;
}

// Address part of one probe = all op-mutants of node e whose edit falls into digit group g: ONE bucket.
struct SlrProbe {
    uint32_t rest, bucket, tag;
};
SLR_HD SlrProbe slr_probe_addr(const SlrTableDev &t, uint32_t cs, uint32_t cbase, int g, int op)
{
    const uint32_t lomask = 0x00FFFFFFu >> (8 * g);            // digit groups below g
    const uint32_t himask = ~(0xFFFFFFFFu >> (8 * g));         // digit groups above g (g = 0: none)
    const uint32_t lo = (op == 0) ? cs : ((op == 1) ? (cs >> 2) : ((cs << 2) | cbase));
    SlrProbe pr;
    pr.rest = ((cs & himask) >> 8) | (lo & lomask);
    const uint32_t m = slr_mix24(pr.rest);
    const int tb = 24 - t.bbits;
    pr.bucket = m >> tb;
    pr.tag = m & ((1u << tb) - 1u);
    return pr;
}

// Overflow stash of a full bucket (cold: P(bucket overflows) ~ 1e-4 for random lists).  Out of line and called with
// scalars only: a struct handed over by reference would have to live in local memory in the hot loop as well.
// Returns best << 32 | barcode.
#define SLR_COLD static __host__ __device__ __noinline__
SLR_COLD unsigned long long slr_probe_stash(const uint32_t *st_bucket, const uint16_t *st_slot, int st_total,
                                            const unsigned long long *vh, uint32_t cs, uint32_t w, uint32_t epack, uint32_t gop,
                                            uint32_t rest, uint32_t bucket, uint32_t tag, uint32_t best, uint32_t bc)
{
    SlrExpand e;
    e.cs = cs; e.w = w; e.pskip = (int)(epack & 31u) - 1; e.cbase = (epack >> 5) & 3u; e.tproc = (epack >> 8) & 0xFFu;
    e.level = (int)((epack >> 16) & 3u); e.use_visited = (epack >> 18) & 1u; e.nopost = (epack >> 19) & 1u;
    const int g = (int)(gop >> 2), op = (int)(gop & 3u);
    const uint32_t want_hi = 0x80u | tag, gb = ((uint32_t)g << 24) | bucket;
    for (int i = slr_stash_lower(st_bucket, st_total, gb); i < st_total && slr_ldg(st_bucket + i) == gb; i++) {
        const uint32_t sl = slr_ldg(st_slot + i);
        if ((sl >> 8) != want_hi) continue;
        uint32_t s = 0;
        const uint32_t r = slr_check_pattern(e, vh, g, op, sl & 0xFFu, rest, s);
        if (r < best) { best = r; bc = s; }
    }
    return ((unsigned long long)best << 32) | bc;
}

// Evaluate a loaded bucket: best (smallest) traversal rank and the matching barcode.
SLR_HD uint32_t slr_probe_eval(const SlrTableDev &t, const SlrExpand &e, const unsigned long long *vh, int g, int op,
                                                   const SlrProbe &pr, const SlrBucket &k, uint32_t &bc_out)
{
    // Tag filter AND a byte-parallel necessary condition on the pattern, the same straight-line code for all three ops
    // (lanes of one warp mix them): a one-edit pattern agrees with the node's digit group either in the two leading
    // digits or in two trailing digits (SUB: the other half; INS: digits 2,3 = the node's 1,2; DEL: digits 1,2 = the
    // node's 2,3 and digit 3 = the digit that follows the group).  1/8 (DEL 1/32) of the unrelated slots survive.
    const uint32_t csg0 = (e.cs >> (24 - 8 * g)) & 0xFFu;
    const uint32_t nx = (g < 3) ? ((e.cs >> (22 - 8 * g)) & 3u) : e.cbase;
    const bool fix3 = op == 2 && !(e.nopost && g == 3);          // a deletion pins digit 3 (unless every appended base is tried)
    const uint32_t want = (0x80u | pr.tag) * 0x01010101u;
#if SLR_NIBBLE_FILTER
    // the two alternatives are kept nibble-disjoint — leading digits (high nibble) as the node's, or trailing digits (low nibble) as the op
    // shifts them: SUB the node's, INS the node's digits 1,2, DEL the node's digit 3 and the digit after the group — so that ONE masked XOR and
    // one zero-NIBBLE test per word answer both (a DEL keeps only its last two digits of the old three-digit test: weaker, still necessary)
    const uint32_t mB = (op == 2 && !fix3) ? 0x0Cu : 0x0Fu;
    const uint32_t vB = (op == 0) ? (csg0 & 0x0Fu) : ((op == 1) ? ((csg0 >> 2) & 0x0Fu) : (((csg0 & 3u) << 2) | (fix3 ? nx : 0u)));
    const uint32_t M4 = (0xF0u | mB) * 0x01010101u, V4 = ((csg0 & 0xF0u) | vB) * 0x01010101u;
#define SLR_PAT_OK(bw, tw) (slr_nibble_ok(((bw) ^ V4) & M4) & slr_eq_bytes((tw), want))
    uint32_t match = (SLR_PAT_OK(k.b.x, k.a.x) >> 7) | (SLR_PAT_OK(k.b.y, k.a.y) >> 6) | (SLR_PAT_OK(k.b.z, k.a.z) >> 5) | (SLR_PAT_OK(k.b.w, k.a.w) >> 4);
#undef SLR_PAT_OK
#else
    const uint32_t mA = fix3 ? 0xF3u : 0xF0u, vA = (csg0 & 0xF0u) | (fix3 ? nx : 0u);
    const uint32_t mB = (op == 2) ? (fix3 ? 0x3Fu : 0x3Cu) : 0x0Fu;
    const uint32_t vB = (op == 0) ? (csg0 & 0x0Fu) : ((op == 1) ? ((csg0 >> 2) & 0x0Fu) : (((csg0 & 0x0Fu) << 2) | (fix3 ? nx : 0u)));
    const uint32_t mA4 = mA * 0x01010101u, vA4 = vA * 0x01010101u, mB4 = mB * 0x01010101u, vB4 = vB * 0x01010101u;
    uint32_t match = ((slr_eq_bytes(k.a.x, want) & (slr_eq_bytes(k.b.x & mA4, vA4) | slr_eq_bytes(k.b.x & mB4, vB4))) >> 7) |
                     ((slr_eq_bytes(k.a.y, want) & (slr_eq_bytes(k.b.y & mA4, vA4) | slr_eq_bytes(k.b.y & mB4, vB4))) >> 6) |
                     ((slr_eq_bytes(k.a.z, want) & (slr_eq_bytes(k.b.z & mA4, vA4) | slr_eq_bytes(k.b.z & mB4, vB4))) >> 5) |
                     ((slr_eq_bytes(k.a.w, want) & (slr_eq_bytes(k.b.w & mA4, vA4) | slr_eq_bytes(k.b.w & mB4, vB4))) >> 4);
#endif
    uint32_t best = SLR_NONE32;
    while (match) {
        const int b = slr_ffs(match) - 1;
        match &= match - 1u;
        const uint32_t P = slr_bucket_pat(k, b);
        uint32_t s = 0;
        const uint32_t r = slr_check_pattern(e, vh, g, op, P, pr.rest, s);
        if (r < best) { best = r; bc_out = s; }
    }
    if (t.st_total > 0 && slr_bucket_full(k)) {                // overflowed bucket: rare, kept out of line
        const uint32_t epack = (uint32_t)(e.pskip + 1) | (e.cbase << 5) | (e.tproc << 8) | ((uint32_t)e.level << 16) |
                               ((e.use_visited ? 1u : 0u) << 18) | ((e.nopost ? 1u : 0u) << 19);
        const unsigned long long r = slr_probe_stash(t.st_bucket, t.st_slot, t.st_total, vh, e.cs, e.w, epack, (uint32_t)(g * 4 + op),
                                                     pr.rest, pr.bucket, pr.tag, best, bc_out);
        best = (uint32_t)(r >> 32);
        bc_out = (uint32_t)r;
    }
    return best;
}

SLR_HD uint32_t slr_expand_group(const SlrTableDev &t, const SlrExpand &e, const unsigned long long *vh, int g, int op,
                                                     uint32_t &bc_out)
{
    const SlrProbe pr = slr_probe_addr(t, e.cs, e.cbase, g, op);
    const SlrBucket k = slr_load_bucket(t, g, pr.bucket);
    return slr_probe_eval(t, e, vh, g, op, pr, k, bc_out);
}

// counters the Java attaches to a mutant created by idx (0-3 SUB -> nSubstitutions, 4-7 INS -> nDeletions (L290),
// 8 DEL -> nInsertions (L346)); packed nSub | nIns << 2 | nDel << 4
SLR_HD uint32_t slr_cnt_of(uint32_t idx) { return idx < 4 ? 1u : (idx < 8 ? (1u << 4) : (1u << 2)); }

// A level-1 node the reference expands: packed for the level-2 loop.
//   meta bits 0-7 = processing time p*16 + jj (jj = 8 - j: nodes of one root position are popped in reverse
//   creation order, ArrayDeque add / pollLast, L212-L218), bits 8-9 = cbase of the node's deletions
//   (post[nDel+1]; nDel = 1 below an INS node), bits 10-15 = counters of the level-1 edit (slr_cnt_of).
SLR_HD uint32_t slr_node_meta(int p, int j, uint32_t p1, uint32_t p2)
{
    const uint32_t cb = (j >= 4 && j < 8) ? p2 : p1;
    return (uint32_t)(p * 16 + (8 - j)) | (cb << 8) | (slr_cnt_of((uint32_t)j) << 10);
}
SLR_HD SlrExpand slr_root_expand(uint32_t w, uint32_t p1, bool use_visited)
{
    SlrExpand e;
    e.cs = w; e.w = w; e.pskip = -1; e.cbase = p1; e.tproc = 0; e.level = 1; e.use_visited = use_visited; e.nopost = false;
    return e;
}
SLR_HD SlrExpand slr_node_expand(uint32_t cs, uint32_t meta, uint32_t w)
{
    SlrExpand e;
    e.cs = cs; e.w = w; e.pskip = (int)((meta >> 4) & 15u); e.cbase = (meta >> 8) & 3u; e.tproc = meta & 0xFFu; e.level = 2; e.use_visited = true; e.nopost = false;
    return e;
}

// ---- character classes -----------------------------------------------------------------------------------
// BASE_TO_TWOBIT_ARRAY (T!…NucleicAcidTwoBitPerBase.java:L78-L87): returns 0..3 or 4 for "not ACGT"
SLR_HD uint32_t slr_code2(uint32_t c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'G': case 'g': return 1; case 'C': case 'c': return 2; case 'T': case 't': return 3;
    default: return 4;
    }
}
// is c in NucleicAcidByteCodeBase.ENCODE_MATRIX (T!…NucleicAcidByteCodeBase.java:L45-L78)?
SLR_HD bool slr_in_encode_matrix(uint32_t c)
{
    if (c == '-') return true;
    switch (c | 0x20u) {
    case 'a': case 'g': case 'c': case 't': case 'n': case 'h': case 'r': case 'y': case 'm': case 'k': case 's':
    case 'w': case 'b': case 'v': case 'd':
        return true;                         // only letters fold onto 'a'..'z' under |0x20
    default: return false;
    }
}

// reverse the order of the 16 base-4 digits of x
SLR_HD uint32_t slr_rev_digits(uint32_t x)
{
    x = slr_brev(x);
    return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}
// spread the low 16 bits of x to the even bit positions
SLR_HD uint32_t slr_spread16(uint32_t x)
{
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// Per-slice bit planes: bit i describes char i of the slice (on the GPU these are warp ballots).
struct SlrSliceBits {
    uint32_t bit0, bit1;      // 2-bit code planes
    uint32_t nonacgt;         // char is not ACGTacgt
    uint32_t unknown;         // char is not in ENCODE_MATRIX
    uint32_t over253;         // char >= 254 (indexing past the 254-entry tables -> AIOOBE)
};

// Window + post bases of offset o.  Returns false if the Java would have thrown (Parser.java:L214-L219).
// w = OneMatch.readSeq; dead = window can never match (5' window with a non-ACGT char: bits >= 32 set).
// need_post = false: the pass-1 exact lookup (UsedCellBCListGenerator$Worker, UsedCellBCListGenerator.java:L210-L221) takes
// the window only, no post sequence.
SLR_HD bool slr_window(const SlrSliceBits &b, int len, int anc, int o, int three_prime, int ed_max, uint32_t &w, uint32_t &p1,
                       uint32_t &p2, bool &dead, bool need_post = true)
{
    const int ws = anc + o;
    dead = false;
    p1 = p2 = 0;
    if (ws < 0 || ws + 16 > len) return false;                              // substring(bcStart-1, bcEnd)
    if (!need_post) {
    } else if (three_prime) {
        if (ws - 4 < 0) return false;                                       // substring(bcStart-5, bcStart)
        if ((b.unknown >> (ws - 4)) & 0x1Fu) return false;                  // reverseComplement(): ONEBYTE_REVERSECOMP_MATRIX[-1]
        // post[1] = comp(read[ws]), post[2] = comp(read[ws-1]); codes other than A,G,C,T append A (BYTE_TO_2BITLONG_ARRAY)
        const uint32_t a1 = ((b.bit1 >> ws) & 1u) * 2u + ((b.bit0 >> ws) & 1u);
        const uint32_t a2 = ((b.bit1 >> (ws - 1)) & 1u) * 2u + ((b.bit0 >> (ws - 1)) & 1u);
        p1 = ((b.nonacgt >> ws) & 1u) ? 0u : 3u - a1;
        p2 = ((b.nonacgt >> (ws - 1)) & 1u) ? 0u : 3u - a2;
    } else {
        if (ws + 21 > len) return false;                                    // substring(bcEnd, bcEnd+5)
        if ((b.over253 >> (ws + 16)) & 0x1Fu) return false;                 // ENCODE_MATRIX[c], 254 entries
        // a code of -1 only throws when a deletion fetches it: post[1] at ED >= 1, post[2] at ED >= 2
        if (ed_max >= 1 && ((b.unknown >> (ws + 16)) & 1u)) return false;
        if (ed_max >= 2 && ((b.unknown >> (ws + 17)) & 1u)) return false;
        p1 = ((b.nonacgt >> (ws + 16)) & 1u) ? 0u : ((b.bit1 >> (ws + 16)) & 1u) * 2u + ((b.bit0 >> (ws + 16)) & 1u);
        p2 = ((b.nonacgt >> (ws + 17)) & 1u) ? 0u : ((b.bit1 >> (ws + 17)) & 1u) * 2u + ((b.bit0 >> (ws + 17)) & 1u);
    }
    if ((b.over253 >> ws) & 0xFFFFu) return false;                          // BASE_TO_TWOBIT_ARRAY[c], 254 entries
    // getLongHashForSeq (java:L183-L187): a non-ACGT char ORs a sign-extended (byte)-2 into the hash: every
    // earlier digit becomes 3, that digit 2, bits >= 32 garbage.  wbits holds char ws+i at digit i from the LSB.
    uint32_t wbits = slr_spread16(b.bit0 >> ws) | (slr_spread16(b.bit1 >> ws) << 1);
    const uint32_t badw = (b.nonacgt >> ws) & 0xFFFFu;
    if (badw) {
        const int jlast = 31 - slr_clz(badw);
        wbits |= slr_lowmask(2 * jlast);
        wbits = (wbits & ~(3u << (2 * jlast))) | (2u << (2 * jlast));
        dead = !three_prime;
    }
    // 3': reverseComplement (java:L477-L484) reads only the low 32 bits; wbits is already reversed, complement = ~
    w = three_prime ? ~wbits : slr_rev_digits(wbits);
    return true;
}

// ---- merge + decision (Parser.java:L240-L311) -----------------------------------------------------------
// Emulates the iteration order of the merged java.util.HashSet<OneMatch> (hashCode = (int)(readSeq ^ readSeq>>>32),
// BarcodeMatchTester.java:L443; JDK HashMap: capacity 16, resize above 12/24/48 entries, a chain reaching 9 nodes
// resizes while capacity < 64), then the stable sort by OneMatch.compareTo (L449-L461), distinctByKey(matchingBC).
// Fills everything but rank (needs the index map) and returns the best barcode's ED level, or -1.
// Capacity of the merged HashMap after all insertions (cold: only reached with >= 9 entries, see slr_decide).
SLR_COLD int slr_decide_cap(const SlrMatchStore &S, int noff, bool &treeified)
{
    int cap = 16, thr = 12, cnt = 0;
    treeified = false;
#pragma unroll 1
    for (int k = 0; k < noff; k++) {
        const uint32_t sp = S.m_w[k] ^ (S.m_w[k] >> 16);
#pragma unroll 1
        for (int lv = 0; lv < 3; lv++) {
            if (!((S.m_valid[k] >> lv) & 1)) continue;
            int chain = slr_popc(S.m_valid[k] & ((1u << lv) - 1u));
#pragma unroll 1
            for (int k2 = 0; k2 < k; k2++) {
                const uint32_t sp2 = S.m_w[k2] ^ (S.m_w[k2] >> 16);
                if (((sp2 ^ sp) & (uint32_t)(cap - 1)) == 0u) chain += slr_popc(S.m_valid[k2] & 7u);
            }
            cnt++;
            if (chain >= 8) { if (cap < 64) { cap <<= 1; thr <<= 1; } else treeified = true; }
            if (cnt > thr) { cap <<= 1; thr <<= 1; }
        }
    }
    return cap;
}

// Which ED-2 searches can still change the record once levels 0 and 1 of every window are known?  (The reference runs
// all of them; the kernel skips the ones whose result provably cannot reach the output.)
//   SLR_L2_ALL   every window (too many entries to pin the HashSet capacity)
//   SLR_L2_TWO   nothing matched at ED <= 1: windows in order until two DIFFERENT barcodes have been hit at ED 2 - the best
//                match then has ED 2 and so has the second best, i.e. the read is unassigned with ed = ed_second = 2
//                whatever the remaining windows add (an unassigned record carries nothing else)
//   SLR_L2_NONE  none: the best match has ED <= 1 and another barcode already sits at ED <= 1, so neither best nor
//                ed_second can move
//   SLR_L2_UNTIL windows in order until one yields an ED-2 hit on a barcode other than bcA: every ED <= 1 entry
//                carries bcA, so best is fixed and ed_second is 2 if such a hit exists, else none
// With n01 + noff <= 8 entries in total the merged HashMap stays at capacity 16 whatever the ED-2 searches add (no
// resize above 12, no bin of 9), so the tie order among the ED <= 1 entries - and with it the best match - is final.
enum { SLR_L2_NONE = 0, SLR_L2_ALL = 1, SLR_L2_UNTIL = 2, SLR_L2_TWO = 3 };
SLR_HD int slr_level2_plan(const SlrMatchStore &S, int noff, uint32_t &bcA)
{
    int n01 = 0;
#pragma unroll 1
    for (int k = 0; k < noff; k++) n01 += slr_popc(S.m_valid[k] & 3u);
    bcA = 0;
    if (n01 + noff > 8) return SLR_L2_ALL;
    if (n01 == 0) return SLR_L2_TWO;
    uint32_t bestkey = SLR_NONE32;
#pragma unroll 1
    for (int k = 0; k < noff; k++) {
        const uint32_t sp = S.m_w[k] ^ (S.m_w[k] >> 16);
#pragma unroll
        for (int lv = 0; lv < 2; lv++) {
            if (!((S.m_valid[k] >> lv) & 1)) continue;
            const uint32_t key = ((uint32_t)lv << 20) | ((k != 0 ? 1u : 0u) << 16) | ((sp & 15u) << 8) | (uint32_t)(k * 3 + lv);
            if (key < bestkey) { bestkey = key; bcA = S.m_bc[k][lv]; }
        }
    }
    bool other = false;
#pragma unroll 1
    for (int k = 0; k < noff; k++)
#pragma unroll
        for (int lv = 0; lv < 2; lv++)
            if (((S.m_valid[k] >> lv) & 1) && S.m_bc[k][lv] != bcA) other = true;
    return other ? SLR_L2_NONE : SLR_L2_UNTIL;
}

SLR_HD int slr_decide(const SlrMatchStore &S, int noff, int ed_max, slr_bc_result &res)
{
    int cnt = 0;
#pragma unroll 1
    for (int k = 0; k < noff; k++) cnt += slr_popc(S.m_valid[k] & 7u);
    int cap = 16;                                  // <= 8 entries: no bin reaches 9 nodes, no resize above 12
    bool treeified = false;
    if (cnt > 8) cap = slr_decide_cap(S, noff, treeified);
    if (cnt == 0) return -1;
    if (treeified) res.flags |= SLR_F_TIE_UNPIN;
    uint32_t bestkey = SLR_NONE32;
    int bk = 0, blv = 0;
#pragma unroll 1
    for (int k = 0; k < noff; k++) {
        const uint32_t sp = S.m_w[k] ^ (S.m_w[k] >> 16);
#pragma unroll
        for (int lv = 0; lv < 3; lv++) {
            if (!((S.m_valid[k] >> lv) & 1)) continue;
            const uint32_t key = ((uint32_t)lv << 20) | ((k != 0 ? 1u : 0u) << 16) | ((sp & (uint32_t)(cap - 1)) << 8) | (uint32_t)(k * 3 + lv);
            if (key < bestkey) { bestkey = key; bk = k; blv = lv; }
        }
    }
    const uint32_t bbc = S.m_bc[bk][blv];
    int second = 0x7FFFFFFF;
#pragma unroll 1
    for (int k = 0; k < noff; k++)
#pragma unroll
        for (int lv = 0; lv < 3; lv++)
            if (((S.m_valid[k] >> lv) & 1) && S.m_bc[k][lv] != bbc && lv < second) second = lv;
    res.ed = blv;
    res.ed_second = second;
    if (blv <= ed_max && blv < second) {                                    // L251-L252
        res.flags |= SLR_F_ASSIGNED;
        res.bc = bbc;
        res.offset = (int8_t)slr_offset_of(bk);
        const uint32_t c = S.m_cnt[bk][blv];
        res.n_sub = (int8_t)(c & 3u);
        res.n_ins = (int8_t)((c >> 2) & 3u);
        res.n_del = (int8_t)((c >> 4) & 3u);
        return blv;
    }
    return -1;
}
