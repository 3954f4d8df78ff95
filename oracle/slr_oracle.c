/*
 * slr_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See slr_oracle.h for the parity status (pinned against the
 * reference's own class files executed by oracle/minijvm.py; what is only restated is listed there).
 *
 * Every function restates one reference method, cited as  jar!class (File.java:Lnnn).
 *   F! = /root/reference/Jar/NanoporeBC_UMI_finder-2.1.jar
 *   T! = /root/reference/Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar
 */
#include "slr_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Java `long << n` / `>>> n` use only the low 6 bits of n (JLS 15.19). */
static inline uint64_t jshl(uint64_t x, int n)  { return x << (n & 63); }
static inline uint64_t jushr(uint64_t x, int n) { return x >> (n & 63); }

/* ------------------------------------------------------------------------------------------------
 * T!com/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase
 * ---------------------------------------------------------------------------------------------- */

/* BASE_TO_TWOBIT_ARRAY (java:L78-L87): 254 entries pre-filled with (byte)-2, A/a=0 G/g=1 C/c=2 T/t=3 */
static inline int base_to_twobit(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'G': case 'g': return 1;
    case 'C': case 'c': return 2;
    case 'T': case 't': return 3;
    default: return -2;
    }
}

/* getLongHashForSeq (java:L183-L187): retval = (retval << 2) | (long) BASE_TO_TWOBIT_ARRAY[c].
 * A non-ACGT char ORs the sign-extended (byte)-2 = 0xFFFF...FE into the hash.  chars >= 254 index past
 * the 254-entry table -> ArrayIndexOutOfBoundsException (*bad_char = 1). */
uint64_t orc_pack2bit(const uint8_t *chars, int len, int *bad_char)
{
    uint64_t h = 0;
    for (int i = 0; i < len; i++) {
        if (chars[i] >= 254) { if (bad_char) *bad_char = 1; }
        h = (h << 2) | (uint64_t)(int64_t)base_to_twobit(chars[i]);
    }
    return h;
}

/* reverseComplement (java:L477-L484), REVERSE_COMP_ARRAY 0<->3, 1<->2 (java:L72-L76).
 * Consumes only the low 2*len bits of the source: garbage high bits are dropped here. */
uint64_t orc_revcomp2bit(uint64_t seq, int len)
{
    uint64_t t = 0;
    for (int i = 0; i < len; i++) {
        t = (t << 2) | (uint64_t)(3 - (seq & 3));
        seq >>= 2;
    }
    return t;
}

/* getLongHashReplaceByteDeg (java:L228-L234) with get0bitsForPosition (L146) = CLEAR_BITS[len-pos-1] */
void orc_replace_deg(uint64_t seq, uint64_t out[4], int pos, int len)
{
    int k = len - pos - 1;                         /* CLEAR_BITS_TWOBIT_ARRAY[k] clears bits 2k,2k+1 (L89-L100) */
    uint64_t clear = ~jshl(3ULL, 2 * k);
    seq &= clear;
    int shift = (len - (pos + 1)) << 1;
    out[0] = seq | jshl(0ULL, shift);
    out[1] = seq | jshl(1ULL, shift);
    out[2] = seq | jshl(2ULL, shift);
    out[3] = seq | jshl(3ULL, shift);
}

/* getLongHashInsertByteDeg (java:L300-L310): insert a base AFTER pos, drop the last base.
 * SET_BITS_TWOBIT_ARRAY[i] = {0,1,3,2} << 2i (L105-L112) indexed with b = 0,1,3,2 -> values 0,1,2,3.
 * Shift-overflow kept: pos == len-2 gives shift = 62 and `>>> 64` == `>>> 0`. */
void orc_insert_deg(uint64_t hash, uint64_t out[4], int pos, int len)
{
    int shift = (len - pos - 1) << 1;
    uint64_t upper = jshl(jushr(hash, shift), shift);
    shift = 64 - shift;
    hash = jshl(hash, shift);
    hash = jushr(hash, shift + 2);
    int i = len - (pos + 1) - 1;                   /* getSeqbitsForPosition(pos+1, len, b) row index */
    for (int v = 0; v < 4; v++)
        out[v] = upper | hash | jshl((uint64_t)v, 2 * i);
}

/* getLongHashdeleteByte (java:L321-L327); BYTE_TO_2BITLONG_ARRAY[0][code] is non-zero only for the
 * 4-bit codes of G(2)->1, C(4)->2, T(8)->3 (L92-L98); everything else (incl. N=15) appends A. */
uint64_t orc_delete_byte(uint64_t hash, int code4, int pos, int len)
{
    int shift = (len - pos) << 1;
    uint64_t upper = jshl(jushr(hash, shift), shift);
    shift = 64 - shift;
    hash = jshl(hash, shift + 2);
    hash = jushr(hash, shift);
    uint64_t add = code4 == 2 ? 1 : code4 == 4 ? 2 : code4 == 8 ? 3 : 0;
    return upper | hash | add;
}

/* T!com/rw/nuc/encoding/NucleicAcidByteCodeBase.ENCODE_MATRIX (java:L45-L78): 254 entries, fill -1 */
int orc_encode4bit(uint8_t c)
{
    switch (c) {
    case '-': return 0;
    case 'A': case 'a': return 1;
    case 'G': case 'g': return 2;
    case 'C': case 'c': return 4;
    case 'T': case 't': return 8;
    case 'N': case 'n': return 15;
    case 'H': case 'h': return 13;
    case 'R': case 'r': return 3;
    case 'Y': case 'y': return 12;
    case 'M': case 'm': return 5;
    case 'K': case 'k': return 10;
    case 'S': case 's': return 6;
    case 'W': case 'w': return 9;
    case 'B': case 'b': return 14;
    case 'V': case 'v': return 7;
    case 'D': case 'd': return 11;
    default: return 0xFF;
    }
}

/* ONEBYTE_REVERSECOMP_MATRIX (java:L100-L133): complement of the IUPAC set = swap A<->T and G<->C bits */
int orc_revcomp4bit(int code)
{
    if (code < 0 || code > 15) return 0xFF;
    return ((code & 1) << 3) | ((code & 8) >> 3) | ((code & 2) << 1) | ((code & 4) >> 1);
}

/* ------------------------------------------------------------------------------------------------
 * search set (membership only; reference: fastutil LongSet view of Long2ObjectOpenHashMap, Parser.java:L228)
 * ---------------------------------------------------------------------------------------------- */
struct orc_set {
    uint64_t *slot_key;
    int32_t  *slot_idx;      /* -1 = empty */
    uint64_t  mask;
    int64_t   n;
};

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

orc_set *orc_set_new(const uint64_t *keys, int64_t n)
{
    orc_set *s = (orc_set *)calloc(1, sizeof(*s));
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    s->mask = cap - 1;
    s->n = 0;
    s->slot_key = (uint64_t *)malloc(cap * sizeof(uint64_t));
    s->slot_idx = (int32_t *)malloc(cap * sizeof(int32_t));
    for (uint64_t i = 0; i < cap; i++) s->slot_idx[i] = -1;
    for (int64_t i = 0; i < n; i++) {
        uint64_t h = mix64(keys[i]) & s->mask;
        int dup = 0;
        while (s->slot_idx[h] >= 0) {
            if (s->slot_key[h] == keys[i]) { dup = 1; break; }
            h = (h + 1) & s->mask;
        }
        if (dup) continue;                         /* first index wins, like Map.put on a fresh map */
        s->slot_key[h] = keys[i];
        s->slot_idx[h] = (int32_t)i;
        s->n++;
    }
    return s;
}

void orc_set_free(orc_set *s)
{
    if (!s) return;
    free(s->slot_key); free(s->slot_idx); free(s);
}

int64_t orc_set_find(const orc_set *s, uint64_t key)
{
    uint64_t h = mix64(key) & s->mask;
    while (s->slot_idx[h] >= 0) {
        if (s->slot_key[h] == key) return s->slot_idx[h];
        h = (h + 1) & s->mask;
    }
    return -1;
}

int64_t orc_set_size(const orc_set *s) { return s->n; }

/* ------------------------------------------------------------------------------------------------
 * visited set: F!com/rw/nuc/encoding/TwoBit/ed/NucTwoBitPerBaseEDtesterBase (java:L82-L120)
 *   active only when ed >= 2;  len > 16 -> LongHashSet(seq), else IntHashSet((int) seq)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *keys; uint8_t *used; uint64_t mask; uint64_t count;
    int active; int use64;
} vset;

static void vset_init(vset *v, int ed, int len)
{
    memset(v, 0, sizeof(*v));
    v->active = ed >= 2;                           /* MINED_TOHASHTESTED_* = 2 (L82-L83) */
    v->use64 = len > 16;                           /* L82: len>=14 && len>16 -> 64 bit; else int */
    if (!v->active) return;
    uint64_t cap = 512;
    v->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    v->used = (uint8_t *)calloc(cap, 1);
    v->mask = cap - 1;
}
static uint64_t g_max_visited;                 /* statistics for sizing the GPU tables (not thread safe: single-threaded probes only) */
uint64_t orc_debug_max_visited(int reset) { uint64_t r = g_max_visited; if (reset) g_max_visited = 0; return r; }
static void vset_free(vset *v) { if (v->count > g_max_visited) g_max_visited = v->count; free(v->keys); free(v->used); }
static inline uint64_t vset_key(const vset *v, uint64_t seq) { return v->use64 ? seq : (uint64_t)(uint32_t)seq; } /* l2i */
static int vset_contains(const vset *v, uint64_t seq)          /* checkWhetherAlreadyTested (L120) */
{
    if (!v->active) return 0;
    uint64_t k = vset_key(v, seq), h = mix64(k) & v->mask;
    while (v->used[h]) { if (v->keys[h] == k) return 1; h = (h + 1) & v->mask; }
    return 0;
}
static void vset_add(vset *v, uint64_t seq)                    /* addToTestedSeqs (L105-L112) */
{
    if (!v->active) return;
    if ((v->count + 1) * 2 > v->mask + 1) {
        uint64_t ocap = v->mask + 1, ncap = ocap * 2;
        uint64_t *ok = v->keys; uint8_t *ou = v->used;
        v->keys = (uint64_t *)malloc(ncap * sizeof(uint64_t));
        v->used = (uint8_t *)calloc(ncap, 1);
        v->mask = ncap - 1;
        for (uint64_t i = 0; i < ocap; i++) if (ou[i]) {
            uint64_t h = mix64(ok[i]) & v->mask;
            while (v->used[h]) h = (h + 1) & v->mask;
            v->used[h] = 1; v->keys[h] = ok[i];
        }
        free(ok); free(ou);
    }
    uint64_t k = vset_key(v, seq), h = mix64(k) & v->mask;
    while (v->used[h]) { if (v->keys[h] == k) return; h = (h + 1) & v->mask; }
    v->used[h] = 1; v->keys[h] = k; v->count++;
}

/* ------------------------------------------------------------------------------------------------
 * F!com/rw/nuc/encoding/TwoBit/LongSeqMutated (java:L44-L77) + NucTwoBitPerBaseWithErrors (L19-L29)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t seq;
    int16_t pos_prev;        /* posTreatedInPreviousLevel */
    int16_t pos_cur;         /* posTreatedInCurrentCycle */
    int16_t level;           /* currentlevel */
    int8_t  n_sub, n_ins, n_del;
} node;

typedef struct {
    const orc_set *set;
    uint64_t unmutated;
    int len, ed, skip_full, allow_indels, do_next, offset;
    const uint8_t *post4; int post_len;
    vset visited;
    node *stack; int sp, scap;                     /* ArrayDeque used LIFO: add / pollLast (L212, L218) */
    orc_match *out; int n_out;
    int64_t probes;
    int exception;
} tester;

static void push(tester *t, const node *n)
{
    if (t->sp == t->scap) {
        t->scap = t->scap ? t->scap * 2 : 64;
        t->stack = (node *)realloc(t->stack, (size_t)t->scap * sizeof(node));
    }
    t->stack[t->sp++] = *n;
}

/* checkMatchWithTestSets (BarcodeMatchTester.java:L367-L374) + Matches.add: HashSet whose equals() is
 * (readSeq, ED, offset) (L433-L436) => within one tester the first hit of every ED level is kept. */
static int check_match(tester *t, const node *n)
{
    if (t->skip_full && t->unmutated == n->seq) return 0;
    t->probes++;
    if (orc_set_find(t->set, n->seq) < 0) return 0;
    for (int i = 0; i < t->n_out; i++)
        if (t->out[i].ed == n->level) return 1;    /* HashSet.add is a no-op, but it WAS a hit */
    orc_match *m = &t->out[t->n_out++];
    m->read_seq = t->unmutated; m->bc = n->seq; m->ed = n->level; m->offset = t->offset;
    m->n_sub = n->n_sub; m->n_ins = n->n_ins; m->n_del = n->n_del;
    return 1;
}

/* goNextEDlevel (NucTwoBitPerBaseEDtesterBase.java:L133-L144); bailoutIfFoundAfterED is null here */
static void go_next(tester *t, const node *n)
{
    if (t->ed > n->level) {
        node c = *n;
        c.pos_prev = n->pos_cur;
        c.pos_cur = -1;
        c.level = (int16_t)(n->level + 1);
        push(t, &c);
    }
}

/* substitutions (BarcodeMatchTester.java:L257-L273) */
static void substitutions(tester *t, const node *cur)
{
    uint64_t mut[4];
    orc_replace_deg(cur->seq, mut, cur->pos_cur, t->len);
    for (int i = 0; i < 4; i++) {
        uint64_t s = mut[i];
        if (s != cur->seq && !vset_contains(&t->visited, s)) {
            node n = *cur;
            n.n_sub++;
            n.seq = s;
            int hit = check_match(t, &n);
            if (hit || t->do_next) go_next(t, &n);           /* L268: rslt != null || doNext */
        }
    }
}

/* insertions (BarcodeMatchTester.java:L284-L300): an inserted base in the candidate == a base deleted
 * from the read, hence the Java bumps nDeletions */
static void insertions(tester *t, const node *cur)
{
    uint64_t mut[4];
    orc_insert_deg(cur->seq, mut, cur->pos_cur, t->len);
    for (int i = 0; i < 4; i++) {
        uint64_t s = mut[i];
        if (!vset_contains(&t->visited, s)) {
            node n = *cur;
            n.seq = s;
            n.n_del++;
            int hit = check_match(t, &n);
            if (!hit || t->do_next) go_next(t, &n);          /* L295: rslt == null || doNext */
        }
    }
}

/* deletions (BarcodeMatchTester.java:L313-L357) */
static void deletions(tester *t, const node *cur)
{
    if (t->post_len >= 0 && cur->n_del + 1 > t->post_len) return;            /* L315 */
    int last = 0;
    if (t->post_len >= 0) {
        last = t->post4[cur->n_del];                                         /* getByteAt(nDel+1), 1-based (L329) */
        if (last > 15) { t->exception = 1; return; }                         /* BYTE_TO_2BITLONG_ARRAY[0][-1] -> AIOOBE */
    }
    uint64_t m = orc_delete_byte(cur->seq, last, cur->pos_cur, t->len);      /* L330 */
    uint64_t cand[4]; int nc;
    if (t->post_len >= 0) { cand[0] = m; nc = 1; }                           /* L332-L334 */
    else { cand[0] = m; cand[1] = m | 1; cand[2] = m | 2; cand[3] = m | 3; nc = 4; }  /* L336-L340 */
    for (int i = 0; i < nc; i++) {
        uint64_t s = cand[i];
        if (!vset_contains(&t->visited, s)) {
            node n = *cur;
            n.seq = s;
            n.n_ins++;
            int hit = check_match(t, &n);
            if (!hit || t->do_next) go_next(t, &n);          /* L351 */
        }
    }
}

/* BarcodeMatchTester.doJob (java:L198-L244) */
int orc_match_tester(const orc_set *set, uint64_t seq, int len, int ed, int skip_full_matches,
                     int allow_indels, const uint8_t *post4, int post_len, int do_next_level_if_match,
                     int offset, orc_match *out, int64_t *n_probes)
{
    tester t;
    memset(&t, 0, sizeof(t));
    t.set = set; t.unmutated = seq; t.len = len; t.ed = ed; t.skip_full = skip_full_matches;
    t.allow_indels = allow_indels; t.do_next = do_next_level_if_match; t.offset = offset;
    t.post4 = post4; t.post_len = post4 ? post_len : -1;
    t.out = out; t.n_out = 0;
    vset_init(&t.visited, ed, len);

    node parent;
    memset(&parent, 0, sizeof(parent));
    parent.seq = seq; parent.pos_prev = -1; parent.pos_cur = -1; parent.level = 0;   /* L198, LongSeqMutated L61 */
    check_match(&t, &parent);                                                        /* L204-L206 */
    if (ed != 0) {
        parent.level = 1;                                                            /* L211 */
        push(&t, &parent);                                                           /* L212 */
        const int len_m1 = len - 1;
        while (t.sp > 0 && !t.exception) {
            node cur = t.stack[--t.sp];                                              /* pollLast (L218) */
            cur.pos_cur++;                                                           /* L222 */
            if (cur.pos_cur < len_m1) push(&t, &cur);                                /* L223-L224 */
            if (cur.pos_prev == cur.pos_cur) continue;                               /* L227-L228 */
            substitutions(&t, &cur);                                                 /* L230 */
            if (allow_indels && cur.pos_cur < len_m1) {                              /* L232-L234 */
                insertions(&t, &cur);                                                /* L235 */
                deletions(&t, &cur);                                                 /* L237 */
            }
            vset_add(&t.visited, cur.seq);                                           /* L241 */
        }
    }
    if (n_probes) *n_probes += t.probes;
    vset_free(&t.visited);
    free(t.stack);
    return t.exception ? -1 : t.n_out;
}

/* ------------------------------------------------------------------------------------------------
 * java.util.HashSet<OneMatch> emulation for the merged `matches` of Parser.assignBarcode (L197, L240):
 * JDK HashMap, default capacity 16 / load factor .75, chains in insertion order, resize keeps relative
 * order; treeifyBin (chain reaching 9 nodes) resizes instead of treeifying while capacity < 64.
 * OneMatch.hashCode = (int)(readSeq ^ readSeq >>> 32) (BarcodeMatchTester.java:L443).
 * ---------------------------------------------------------------------------------------------- */
#define HS_MAX (8 * (ORC_MAX_ED + 1) + 16)
typedef struct {
    orc_match e[HS_MAX];
    uint32_t spread[HS_MAX];
    int n, cap, thr, treeified;
} jhashset;

static void hs_init(jhashset *h) { h->n = 0; h->cap = 16; h->thr = 12; h->treeified = 0; }

static void hs_add(jhashset *h, const orc_match *m)
{
    uint32_t hc = (uint32_t)(m->read_seq ^ (m->read_seq >> 32));
    uint32_t sp = hc ^ (hc >> 16);                                   /* HashMap.hash() */
    uint32_t b = sp & (uint32_t)(h->cap - 1);
    int chain = 0;
    for (int i = 0; i < h->n; i++) {
        if ((h->spread[i] & (uint32_t)(h->cap - 1)) != b) continue;
        chain++;
        if (h->spread[i] == sp && h->e[i].read_seq == m->read_seq && h->e[i].ed == m->ed && h->e[i].offset == m->offset)
            return;                                                  /* equals() -> already present */
    }
    if (h->n >= HS_MAX) return;
    h->e[h->n] = *m; h->spread[h->n] = sp; h->n++;
    if (chain >= 8) {                                                /* binCount >= TREEIFY_THRESHOLD-1 */
        if (h->cap < 64) { h->cap <<= 1; h->thr <<= 1; }             /* treeifyBin -> resize() */
        else h->treeified = 1;
    }
    if (h->n > h->thr) { h->cap <<= 1; h->thr <<= 1; }               /* ++size > threshold -> resize() */
}

/* iteration order: bucket ascending, insertion order inside a bucket */
static int hs_iter(const jhashset *h, int *order)
{
    int k = 0;
    for (int b = 0; b < h->cap; b++)
        for (int i = 0; i < h->n; i++)
            if ((int)(h->spread[i] & (uint32_t)(h->cap - 1)) == b) order[k++] = i;
    return k;
}

/* OneMatch.compareTo (BarcodeMatchTester.java:L449-L461) */
static int match_cmp(const orc_match *a, const orc_match *b)
{
    if (a->ed < b->ed) return -1;
    if (a->ed > b->ed) return 1;
    if (a->offset == 0 && b->offset != 0) return -1;
    if (a->offset != 0 && b->offset == 0) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Parser.assignBarcode (F!com/rw/nanoporereadscanner/analyzers/Parser.class, Parser.java:L195-L315)
 * `slice` is a piece of the stranded read; `anchor` = 0-based index in the slice of the first base of
 * the offset-0 window (3': adapterpos-17, 5': adapterpos, both relative to the slice start).
 * ---------------------------------------------------------------------------------------------- */
void orc_assign_barcode(const orc_set *set, const int32_t *rank, int ed_max, int plusminus, int three_prime,
                        int bc_len, const uint8_t *slice, int slice_len, int anchor,
                        orc_bc_result *out, int64_t *n_probes)
{
    memset(out, 0, sizeof(*out));
    out->ed = -1; out->ed_second = INT32_MAX; out->rank = -1;

    jhashset hs;
    hs_init(&hs);
    const int L = bc_len;
    /* IntStream.rangeClosed(-pm, pm).boxed().sorted(comparingInt(Math::abs)) -> 0,-1,1,-2,2,... (L198-L200) */
    for (int k = 0; k <= 2 * plusminus; k++) {
        int o = (k == 0) ? 0 : ((k & 1) ? -((k + 1) / 2) : (k / 2));
        int ws = anchor + o;                                   /* 0-based window start in the slice */
        if (ws < 0 || ws + L > slice_len) { out->flags |= ORC_F_EXCEPTION; return; }     /* substring (L214) throws */
        uint8_t post4[5];
        if (three_prime) {
            if (ws - 4 < 0) { out->flags |= ORC_F_EXCEPTION; return; }                   /* substring(bcStart-5, bcStart) (L218) */
            /* post = revcomp(read[ws-4 .. ws]) : post[1] = comp(read[ws]) ... post[5] = comp(read[ws-4]) */
            for (int i = 0; i < 5; i++) {
                int c = orc_encode4bit(slice[ws - i]);
                if (c > 15) { out->flags |= ORC_F_EXCEPTION; return; }                   /* ONEBYTE_REVERSECOMP_MATRIX[-1] */
                post4[i] = (uint8_t)orc_revcomp4bit(c);
            }
        } else {
            if (ws + L + 5 > slice_len) { out->flags |= ORC_F_EXCEPTION; return; }       /* substring(bcEnd, bcEnd+5) (L219) */
            for (int i = 0; i < 5; i++) {
                if (slice[ws + L + i] >= 254) { out->flags |= ORC_F_EXCEPTION; return; } /* ENCODE_MATRIX[c] AIOOBE */
                post4[i] = (uint8_t)orc_encode4bit(slice[ws + L + i]);                   /* 0xFF kept: throws only if used */
            }
        }
        int bad = 0;
        uint64_t bc = orc_pack2bit(slice + ws, L, &bad);                                 /* L214 */
        if (bad) { out->flags |= ORC_F_EXCEPTION; return; }
        if (three_prime) bc = orc_revcomp2bit(bc, L);                                    /* L221 */

        orc_match m[ORC_MAX_ED + 1];
        int nm = orc_match_tester(set, bc, L, ed_max, /*skipFullMatches*/0, /*allowIndels*/1,
                                  post4, 5, /*doNextLevelIfMatchFound*/1, o, m, n_probes);   /* L223-L238 */
        if (nm < 0) { out->flags |= ORC_F_EXCEPTION; return; }
        for (int i = 0; i < nm; i++) hs_add(&hs, &m[i]);                                 /* matches.addMatches(m) (L240) */
    }
    if (hs.n == 0) return;                                                               /* L244 */
    if (hs.treeified) out->flags |= ORC_F_TIE_UNPIN;

    /* matches.stream().sorted()  — stable sort of the HashSet iteration order (L247) */
    int order[HS_MAX];
    int n = hs_iter(&hs, order);
    for (int i = 1; i < n; i++) {                        /* insertion sort == stable */
        int x = order[i], j = i - 1;
        while (j >= 0 && match_cmp(&hs.e[order[j]], &hs.e[x]) > 0) { order[j + 1] = order[j]; j--; }
        order[j + 1] = x;
    }
    /* .filter(distinctByKey(matchingBC)).findFirst() / .skip(1).findFirst() (L247-L250) */
    const orc_match *best = &hs.e[order[0]];
    const orc_match *second = NULL;
    for (int i = 1; i < n; i++)
        if (hs.e[order[i]].bc != best->bc) { second = &hs.e[order[i]]; break; }

    out->ed = best->ed;
    out->ed_second = second ? second->ed : INT32_MAX;
    if (best->ed <= ed_max && (second == NULL || best->ed < second->ed)) {               /* L251-L252 */
        out->flags |= ORC_F_ASSIGNED;
        out->bc = best->bc;
        out->offset = (int8_t)best->offset;
        out->n_ins = (int8_t)best->n_ins; out->n_del = (int8_t)best->n_del; out->n_sub = (int8_t)best->n_sub;
        int64_t idx = orc_set_find(set, best->bc);
        out->rank = (rank && idx >= 0) ? rank[idx] : (int32_t)idx;                       /* L267-L269 */
    }
}

void orc_assign_barcode_batch(const orc_set *set, const int32_t *rank, int ed_max, int plusminus, int three_prime,
                              int bc_len, const uint8_t *slices, int stride, int slice_len, const int32_t *anchor,
                              int64_t n, orc_bc_result *out, int64_t *n_probes_total, int n_threads)
{
    int64_t total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total)
#endif
    for (int64_t i = 0; i < n; i++) {
        int64_t p = 0;
        orc_assign_barcode(set, rank, ed_max, plusminus, three_prime, bc_len, slices + i * stride, slice_len,
                           anchor[i], &out[i], &p);
        total += p;
    }
    if (n_probes_total) *n_probes_total = total;
}

/* ------------------------------------------------------------------------------------------------
 * UsedCellBCListGenerator$Worker.call, per read (F!com/rw/nanoporereadscanner/analyzers/UsedCellBCListGenerator$Worker.class,
 * UsedCellBCListGenerator.java:L206-L232): bcStart/bcEnd from the adapter end (L210-L215), bc0 = 2-bit pack of
 * read.substring(bcStart-1, bcEnd) (L218-L219), reverse complement for 3' (L220-L221), then containsKey on the used map /
 * the whitelist predicate (L222-L229).
 * ---------------------------------------------------------------------------------------------- */
void orc_exact_lookup_batch(const orc_set *set, const int32_t *rank, int three_prime, int bc_len, const uint8_t *slices, int stride,
                            int slice_len, const int32_t *lens, const int32_t *anchor, int64_t n, orc_bc_result *out)
{
    for (int64_t i = 0; i < n; i++) {
        orc_bc_result *r = &out[i];
        memset(r, 0, sizeof(*r));
        r->ed = -1; r->ed_second = INT32_MAX; r->rank = -1;
        int len = lens ? (lens[i] < slice_len ? lens[i] : slice_len) : slice_len;
        int ws = anchor[i];
        if (ws < 0 || ws + bc_len > len) { r->flags |= ORC_F_EXCEPTION; continue; }      /* substring throws */
        int bad = 0;
        uint64_t bc = orc_pack2bit(slices + i * (int64_t)stride + ws, bc_len, &bad);
        if (bad) { r->flags |= ORC_F_EXCEPTION; continue; }                              /* BASE_TO_TWOBIT_ARRAY[c >= 254] */
        if (three_prime) bc = orc_revcomp2bit(bc, bc_len);
        int64_t idx = orc_set_find(set, bc);
        if (idx < 0) continue;
        r->flags |= ORC_F_ASSIGNED; r->bc = bc; r->ed = 0;
        r->rank = rank ? rank[idx] : (int32_t)idx;
    }
}

/* ------------------------------------------------------------------------------------------------
 * BarcodeDatasetColissionTester.submitSeq (F!com/rw/nanoporereadscanner/analyzers/BarcodeDatasetColissionTester.class,
 * BarcodeDatasetColissionTester.java:L212-L229): every used barcode is run through the same engine against the used
 * list itself with skipFullMatches = true, allowIndels = true, offset 0, postSeq = null, doNext = false (L215-L222).
 * Matches keeps the first hit of every ED level (HashSet.equals on (readSeq, ED, offset)).
 * ---------------------------------------------------------------------------------------------- */
void orc_collide_batch(const orc_set *set, int ed, int bc_len, const uint64_t *queries, int64_t n, orc_collide_result *out,
                       int64_t *n_probes_total, int n_threads)
{
    int64_t total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total)
#endif
    for (int64_t i = 0; i < n; i++) {
        orc_match m[ORC_MAX_ED + 2];
        int64_t probes = 0;
        orc_collide_result r;
        memset(&r, 0, sizeof(r));
        int k = orc_match_tester(set, queries[i], bc_len, ed, 1, 1, NULL, -1, 0, 0, m, &probes);
        for (int a = 0; a < k; a++) {
            if (m[a].ed < 1 || m[a].ed > 2) continue;
            int lv = m[a].ed - 1;
            r.valid |= (uint8_t)(1u << lv);
            r.bc[lv] = m[a].bc;
            r.n_sub[lv] = (uint8_t)m[a].n_sub; r.n_ins[lv] = (uint8_t)m[a].n_ins; r.n_del[lv] = (uint8_t)m[a].n_del;
        }
        out[i] = r;
        total += probes;
    }
    if (n_probes_total) *n_probes_total += total;
}

/* ------------------------------------------------------------------------------------------------
 * UMI pair distance
 * ---------------------------------------------------------------------------------------------- */

/* F!com/rw/nanopore/analyzers/apachemod/LevenshteinDistance.limitedCompare (java:L220-L283) */
int orc_limited_compare(const uint8_t *left, int n, const uint8_t *right, int m, int threshold)
{
    int p[66], d[66];
    if (n > 64) return -2;
    int *pp = p, *dd = d;
    int boundary = threshold + 1;                                        /* L229 */
    if (boundary > n + 1) boundary = n + 1;                              /* (Java would throw for n < threshold; unused) */
    for (int i = 0; i < boundary; i++) pp[i] = i;                        /* L230-L231 */
    for (int i = boundary; i < n + 1; i++) pp[i] = INT_MAX;              /* L235 */
    for (int i = 0; i < n + 1; i++) dd[i] = INT_MAX;                     /* L236 */
    for (int j = 1; j <= m; j++) {                                       /* L239 */
        uint8_t rj = right[j - 1];
        dd[0] = j;                                                       /* L241 */
        int mn = 1 > j - threshold ? 1 : j - threshold;                  /* L244 */
        int mx = j > INT_MAX - threshold ? n : (n < j + threshold ? n : j + threshold);   /* L245 */
        if (mn > 1) dd[mn - 1] = INT_MAX;                                /* L249-L250 */
        int lower = INT_MAX;                                             /* L253 */
        for (int i = mn; i <= mx; i++) {                                 /* L255 */
            if (left[i - 1] == rj) dd[i] = pp[i - 1];                    /* L256-L258 */
            else {
                int a = dd[i - 1] < pp[i] ? dd[i - 1] : pp[i];
                a = a < pp[i - 1] ? a : pp[i - 1];
                dd[i] = (int)(1u + (unsigned)a);                         /* L262 (int wrap like Java) */
            }
            if (dd[i] < lower) lower = dd[i];                            /* L264 */
        }
        if (lower > threshold) return -1;                                /* L267-L268 */
        int *t = pp; pp = dd; dd = t;                                    /* L272-L274 */
    }
    return pp[n] <= threshold ? pp[n] : -1;                              /* L280-L283 */
}

/* BestEditDistance(ed, pos1, pos2) (ClusteringEditDistanceBase.java:L425-L428); value: MINUSONE=0 ZERO=1 PLUSONE=2 */
static inline int32_t pack_best(int ed, int v1, int v2)
{
    return (int32_t)((ed & 0xFFFFFF) | (0x08000000 << v1) | (0x01000000 << v2));
}

/* calcEditDistances (lambda$static$7, java:L297-L350) + calcBestEditDistance (L67-L80).
 * a,b hold umi_len+2 codes: index 0 is the base BEFORE the predicted UMI start (shift -1). */
int32_t orc_umi_best9(const uint8_t *a, const uint8_t *b, int umi_len)
{
    int8_t eds[3][3];
    for (int i = -1; i < 2; i++)
        for (int j = -1; j < 2; j++) {
            const uint8_t *s1 = a + 1 + i, *s2 = b + 1 + j;             /* getSubSequence(bcEnd+1+i, umi_length) (L322, L329) */
            if (memcmp(s1, s2, (size_t)umi_len) == 0) eds[i + 1][j + 1] = 0;             /* L332-L333 */
            else {
                int d = orc_limited_compare(s1, umi_len, s2, umi_len, 4);                /* L341-L342 */
                eds[i + 1][j + 1] = (int8_t)(d == -1 ? 5 : d);                           /* L343 */
            }
        }
    /* POSITIONS EnumSet iterates in ordinal order ZERO(value 1), PLUSONE(2), MINUSONE(0) (PlusMinusOnePosData.java:L20-L22) */
    static const int order[3] = { 1, 2, 0 };
    int best = 127, b1 = 0, b2 = 0;                                     /* MutableTriple(127, MINUSONE, MINUSONE) (L67) */
    for (int x = 0; x < 3; x++)
        for (int y = 0; y < 3; y++) {
            int i = order[x], v = order[y];
            if (eds[i][v] < best) { best = eds[i][v]; b1 = i; b2 = v; }  /* strict '<' (L73-L76) */
        }
    return pack_best(best, b1, b2);
}

/* BestEditDistance.getTransposedCopy() (java:L458): new BestEditDistance(ed, getPos2(), getPos1()) */
int32_t orc_umi_transpose(int32_t packed)
{
    uint32_t u = (uint32_t)packed;
    uint32_t p1 = (u >> 27) & 7, p2 = (u >> 24) & 7;
    return (int32_t)((u & 0xFFFFFF) | (p2 << 27) | (p1 << 24));
}

/* equalityEditDistance = new ClusteringEditDistanceBase(EQUALITYMATRIX) (java:L90-L92): [[2,1,2],[1,0,1],[2,1,2]] */
int32_t orc_umi_equality(void) { return pack_best(0, 1, 1); }

/* generateDistanceMatrix{NonParallel,Paralell} (java:L168-L259): diagonal = equality (the parallel path
 * recomputes it with calc(i,i), which gives the same (0,ZERO,ZERO)); upper triangle computed, lower = transposed. */
void orc_umi_matrix(const uint8_t *umis, int stride, int umi_len, int64_t n, int32_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        out[i * n + i] = orc_umi_equality();
        for (int64_t v = i + 1; v < n; v++) {
            int32_t e = orc_umi_best9(umis + i * stride, umis + v * stride, umi_len);
            out[i * n + v] = e;
            out[v * n + i] = orc_umi_transpose(e);
        }
    }
}

void orc_umi_matrix_batch(const uint8_t *umis, int stride, int umi_len, const int64_t *job_offsets, int64_t n_jobs,
                          int32_t *out, const int64_t *out_offsets, int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t j = 0; j < n_jobs; j++) {
        int64_t s = job_offsets[j], n = job_offsets[j + 1] - s;
        if (n <= 1500) orc_umi_matrix(umis + s * stride, stride, umi_len, n, out + out_offsets[j]);
    }
    /* giant jobs one at a time, their rows over all threads (a 20 000-read job is 2 x 10^8 pairs: minutes on one core) */
    for (int64_t j = 0; j < n_jobs; j++) {
        const int64_t s = job_offsets[j], n = job_offsets[j + 1] - s;
        if (n <= 1500) continue;
        const uint8_t *u = umis + s * stride;
        int32_t *o = out + out_offsets[j];
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8)
#endif
        for (int64_t i = 0; i < n; i++) {
            o[i * n + i] = orc_umi_equality();
            for (int64_t v = i + 1; v < n; v++) {
                const int32_t e = orc_umi_best9(u + i * stride, u + v * stride, umi_len);
                o[i * n + v] = e;
                o[v * n + i] = orc_umi_transpose(e);
            }
        }
    }
}

/* ================================================================================================
 * Neighbour-set clustering: ClusterOne_MyClustering.clusterLocal (ClusterOne_MyClustering.java:L175-L219)
 *   L179-L185  indices.stream().map(a -> (a, indices.stream().filter(v -> dm.distanceNonReducedSet(a, v) <= ed)
 *              .collect(HashSet))).filter(set.size() > 1).collect(toMap(..., Int2ObjectOpenHashMap::new))
 *   L190-L196  keySet().stream().map(c -> (c, entrySet().stream().filter(l -> l.value.contains(c))
 *              .max(comparing(l -> l.value.size())).get().key))      -- Stream.max keeps the first maximum
 *   L199, L219 grouped by the chosen entry (done by the caller from rec[].best_key)
 * distanceNonReducedSet = matrix[a][v].bestEditDistance.getED() (DistanceMatrix.java:L169), getED = (byte)(ed & 0xFFFFFF).
 * ============================================================================================== */
static int cluster_ed(int32_t packed) { return (int)(int8_t)(packed & 0xFFFFFF); }

void orc_umi_cluster(const int32_t *matrix, int64_t n, int ed, const uint8_t *member, const int32_t *order, int64_t n_order,
                     orc_cluster_rec *rec)
{
    /* possibleClusters: the neighbour sets as one n x n bit matrix (in[a*n+v] = v in N(a)) */
    uint8_t *in = (uint8_t *)calloc((size_t)(n > 0 ? n * n : 1), 1);
    for (int64_t a = 0; a < n; a++) {
        rec[a].n_neighbours = 0; rec[a].best_key = -1; rec[a].best_count = 0; rec[a].n_ties = 0;
        if (member && !member[a]) continue;
        for (int64_t v = 0; v < n; v++) {
            if (member && !member[v]) continue;
            if (cluster_ed(matrix[a * n + v]) <= ed) { in[a * n + v] = 1; rec[a].n_neighbours++; }
        }
    }
    for (int64_t c = 0; c < n; c++) {
        if (rec[c].n_neighbours <= 1) continue;                       /* c is no key */
        const int64_t n_it = order ? n_order : n;
        for (int64_t k = 0; k < n_it; k++) {                          /* entries in iteration order */
            const int64_t l = order ? order[k] : k;
            if (l < 0 || l >= n || rec[l].n_neighbours <= 1) continue;
            if (!in[l * n + c]) continue;                             /* l.getValue().contains(c) */
            if (rec[l].n_neighbours > rec[c].best_count) {            /* compare(a, b) >= 0 ? a : b  -> first maximum stays */
                rec[c].best_count = rec[l].n_neighbours; rec[c].best_key = (int32_t)l; rec[c].n_ties = 1;
            } else if (rec[l].n_neighbours == rec[c].best_count) rec[c].n_ties++;
        }
    }
    free(in);
}

typedef struct { int32_t rank, key; } cluster_rk;
static int cluster_rk_cmp(const void *a, const void *b)
{
    const cluster_rk *x = (const cluster_rk *)a, *y = (const cluster_rk *)b;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    return (x->key > y->key) - (x->key < y->key);
}

void orc_umi_cluster_batch(const int32_t *matrices, const int64_t *job_offsets, const int64_t *out_offsets, int64_t n_jobs, int ed,
                           const uint8_t *member, const int32_t *rank, orc_cluster_rec *rec, int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t j = 0; j < n_jobs; j++) {
        const int64_t s = job_offsets[j], n = job_offsets[j + 1] - s;
        int32_t *order = NULL;
        if (rank && n > 0) {
            cluster_rk *rk = (cluster_rk *)malloc((size_t)n * sizeof(cluster_rk));
            for (int64_t i = 0; i < n; i++) { rk[i].rank = rank[s + i]; rk[i].key = (int32_t)i; }
            qsort(rk, (size_t)n, sizeof(cluster_rk), cluster_rk_cmp);
            order = (int32_t *)malloc((size_t)n * sizeof(int32_t));
            for (int64_t i = 0; i < n; i++) order[i] = rk[i].key;
            free(rk);
        }
        orc_umi_cluster(matrices + out_offsets[j], n, ed, member ? member + s : NULL, order, n, rec + s);
        free(order);
    }
}

/* ================================================================================================
 * Illumina-guided search engine
 * F!com/rw/nuc/encoding/TwoBit/ed/BCUMIEDtesterBase (BCUMIEDtesterBase.java:L82-L203) on top of
 * NucTwoBitPerBaseEDtesterBase (visited set L82-L120, goNextEDlevel with bailout L133-L144).
 * Differences from BarcodeMatchTester.doJob that matter: the root is created with currentlevel 1 (L82,
 * LongSeqMutated.java:L61), every hit is appended to an ArrayList (no first-wins collapse, L88-L89, L145-L146),
 * goNextEDlevel is always called (L148, L172, L201), deletions also run at the last position (L113-L119) and
 * need a non-null post sequence (L187-L190).
 * ============================================================================================== */
typedef struct {
    uint64_t seq;
    int16_t pos_prev, pos_cur, level;
    int8_t  n_sub, n_ins, n_del;
    uint8_t inh;             /* findingErrorFlag of the node: GENE bit, copied by the copy constructor (NucTwoBitPerBaseWithErrors.java:L55) */
} gnode;

typedef struct {
    const orc_guided_sets *sets;
    int len, ed, allow_indels, bailout, offset;
    const uint8_t *post4; int post_len;
    vset visited;
    gnode *stack; int sp, scap;
    orc_guided_hit *out; int64_t cap, n_out;
    int64_t probes;
    int exception;
} gtester;

static void gpush(gtester *t, const gnode *n)
{
    if (t->sp == t->scap) {
        t->scap = t->scap ? t->scap * 2 : 64;
        t->stack = (gnode *)realloc(t->stack, (size_t)t->scap * sizeof(gnode));
    }
    t->stack[t->sp++] = *n;
}

static void glist_add(gtester *t, const gnode *n, unsigned where)
{
    if (t->n_out < t->cap) {
        orc_guided_hit *h = &t->out[t->n_out];
        memset(h, 0, sizeof(*h));
        h->seq = n->seq; h->n_sub = n->n_sub; h->n_ins = n->n_ins; h->n_del = n->n_del;
        h->offset = (int8_t)t->offset; h->where = (uint8_t)where; h->level = (uint8_t)n->level;
    }
    t->n_out++;
}

/* checkMatchWithTestSets + matchingList.add.  UMI flavour (UMInucTwoBitPerBaseEDtester.java:L52-L67, useShortenedUMI is
 * false at the only call site, IlluminaUMIanalyzer.java:L126): umis.contains(seq).  BC flavour
 * (BCnucTwoBitPerBaseEDtester.java:L72-L92): gene set at any level -> the NODE ITSELF gets the GENE bit (L76, inherited by
 * every descendant); else all-passed list while currentlevel <= allPassed10xBCsED (L79-L83); else empty drops while
 * currentlevel <= outOfCellsBarcodesED (L85-L89); the two latter return a COPY (flag stays off the node). */
static void gcheck(gtester *t, gnode *n)
{
    const orc_guided_sets *s = t->sets;
    if (s->group) {
        t->probes++;
        if (orc_set_find(s->group, n->seq) >= 0) {
            if (s->bc_flavour) n->inh = 1;                             /* L76: seq.findingErrorFlag |= GENE */
            glist_add(t, n, n->inh ? ORC_W_GENE : 0u);
            return;
        }
    }
    if (s->all && n->level <= s->all_ed) {
        t->probes++;
        if (orc_set_find(s->all, n->seq) >= 0) { glist_add(t, n, (n->inh ? ORC_W_GENE : 0u) | ORC_W_ALL); return; }
    }
    if (s->empty && n->level <= s->empty_ed) {
        t->probes++;
        if (orc_set_find(s->empty, n->seq) >= 0) { glist_add(t, n, (n->inh ? ORC_W_GENE : 0u) | ORC_W_EMPTY); return; }
    }
}

/* goNextEDlevel (NucTwoBitPerBaseEDtesterBase.java:L133-L144): no push once level >= bailout and the list is non-empty */
static void ggo_next(gtester *t, const gnode *n)
{
    if (t->ed > n->level) {
        if (t->bailout < 0 || n->level < t->bailout || t->n_out == 0) {
            gnode c = *n;
            c.pos_prev = n->pos_cur;
            c.pos_cur = -1;
            c.level = (int16_t)(n->level + 1);
            gpush(t, &c);
        }
    }
}

int64_t orc_guided_tester(const orc_guided_sets *sets, uint64_t seq, int len, int ed, int allow_indels, const uint8_t *post4,
                          int post_len, int bailout, int offset, orc_guided_hit *out, int64_t cap, int64_t *n_probes)
{
    gtester t;
    memset(&t, 0, sizeof(t));
    t.sets = sets; t.len = len; t.ed = ed; t.allow_indels = allow_indels; t.bailout = bailout; t.offset = offset;
    t.post4 = post4; t.post_len = post_len; t.out = out; t.cap = cap;
    vset_init(&t.visited, ed, len);        /* ctor L82-L95: active iff ed >= 2; int keys for len <= 16 */

    gnode parent;
    memset(&parent, 0, sizeof(parent));
    parent.seq = seq; parent.pos_prev = -1; parent.pos_cur = -1; parent.level = 1;     /* L82 */
    gcheck(&t, &parent);                                                                /* L86-L89 */
    if (ed != 0) {                                                                      /* L91-L92 */
        gpush(&t, &parent);                                                             /* L94 */
        const int len_m1 = len - 1;                                                     /* L96 */
        uint64_t mut[4];
        while (t.sp > 0 && !t.exception) {
            gnode cur = t.stack[--t.sp];                                                /* pollLast (L100) */
            cur.pos_cur++;                                                              /* L104 */
            if (cur.pos_cur < len_m1) gpush(&t, &cur);                                  /* L105-L106 */
            if (cur.pos_prev == cur.pos_cur) continue;                                  /* L109-L110 */
            /* substitutions (L136-L151) */
            orc_replace_deg(cur.seq, mut, cur.pos_cur, len);
            for (int i = 0; i < 4; i++) {
                if (mut[i] != cur.seq && !vset_contains(&t.visited, mut[i])) {          /* L138-L139 */
                    gnode n = cur; n.n_sub++; n.seq = mut[i];
                    gcheck(&t, &n);
                    ggo_next(&t, &n);                                                   /* L148 */
                }
            }
            if (allow_indels) {                                                         /* L113 */
                if (cur.pos_cur < len_m1) {                                             /* L115: insertions only */
                    orc_insert_deg(cur.seq, mut, cur.pos_cur, len);
                    for (int i = 0; i < 4; i++) {
                        if (!vset_contains(&t.visited, mut[i])) {                       /* L163 */
                            gnode n = cur; n.seq = mut[i]; n.n_del++;                   /* L165-L166 */
                            gcheck(&t, &n);
                            ggo_next(&t, &n);                                           /* L172 */
                        }
                    }
                }
                /* deletions (L187-L203), also at the last position */
                if (cur.n_del + 1 <= post_len) {                                        /* L187-L188 */
                    int code = post4[cur.n_del];                                        /* getByteAt(nDeletions + 1), 1-based (L190) */
                    if (code > 15) { t.exception = 1; break; }                          /* BYTE_TO_2BITLONG_ARRAY[0][-1] -> AIOOBE */
                    uint64_t m = orc_delete_byte(cur.seq, code, cur.pos_cur, len);
                    if (!vset_contains(&t.visited, m)) {                                /* L192 */
                        gnode n = cur; n.seq = m; n.n_ins++;                            /* L194-L195 */
                        gcheck(&t, &n);
                        ggo_next(&t, &n);                                               /* L201 */
                    }
                }
            }
            vset_add(&t.visited, cur.seq);                                              /* L122 */
        }
    }
    if (n_probes) *n_probes += t.probes;
    vset_free(&t.visited);
    free(t.stack);
    return t.exception ? -1 : t.n_out;
}

/* BarcodeFindingFlag.scoreWhereFound (BarcodeFindingFlag.java:L119-L131); GENE_NOT_FOUND_IN_ILLUMINA is never set by the tester */
static int score_where(unsigned where)
{
    if (where & ORC_W_GENE) return 3;
    if (where & ORC_W_ALL) return 2;
    if (where & ORC_W_EMPTY) return 1;
    return 0;
}

typedef struct { orc_guided_hit h; int64_t idx; int bc; } gsort_item;

/* BCEditDistanceErrorComparator.compare (IlluminaBarcodeUMIAnalyzerBase.java:L106-L110) / EditDistanceErrorComparator
 * (L143-L144); the original index makes qsort reproduce Stream.sorted's stable order */
static int gsort_cmp(const void *pa, const void *pb)
{
    const gsort_item *a = (const gsort_item *)pa, *b = (const gsort_item *)pb;
    int r = (a->h.n_del + a->h.n_ins + a->h.n_sub) - (b->h.n_del + b->h.n_ins + b->h.n_sub);
    if (r == 0 && a->bc) r = score_where(a->h.where) - score_where(b->h.where);
    if (r == 0) r = abs(a->h.offset) - abs(b->h.offset);
    if (r == 0) r = a->idx < b->idx ? -1 : a->idx > b->idx ? 1 : 0;
    return r;
}

void orc_guided_query(const orc_guided_sets *sets, int bc_flavour, int len, int ed, int plusminus, int bailout, int post_len,
                      const uint8_t *slice, int slice_len, int anchor, orc_guided_result *out, orc_guided_hit *raw_out,
                      int64_t raw_cap, int64_t *n_probes)
{
    memset(out, 0, sizeof(*out));
    out->min_err_gene = INT32_MAX;
    int64_t cap = 256, n = 0;
    orc_guided_hit *list = (orc_guided_hit *)malloc((size_t)cap * sizeof(*list));
    /* IntStream.rangeClosed(-pm, pm).boxed().sorted(comparingInt(Math::abs)) (IlluminaUMIanalyzer.java:L89-L91,
     * IlluminaBarcodeAnalyzer.java:L272-L278): 0,-1,1,-2,2 */
    for (int k = 0; k <= 2 * plusminus; k++) {
        const int o = (k == 0) ? 0 : ((k & 1) ? -((k + 1) / 2) : (k / 2));
        const int ws = anchor + o;
        if (ws < 0 || ws + len + post_len > slice_len) goto exception;                  /* getSubSequence out of range */
        uint64_t w = 0;
        for (int i = 0; i < len; i++) {            /* new NucleicAcidTwoBitPerBase(bytes): getLongHashForBytes (T!...java:L197-L201) */
            const int c = orc_encode4bit(slice[ws + i]);
            if (c >= 15) goto exception;           /* unknown char; N = 15 indexes past FOURBIT_TO_TWOBIT_MATRIX[15] (AIOOBE) */
            w = (w << 2) | (uint64_t)(c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 0);       /* other IUPAC codes map to 0 */
        }
        uint8_t post4[32];
        for (int i = 0; i < post_len && i < 32; i++) post4[i] = (uint8_t)orc_encode4bit(slice[ws + len + i]);
        for (;;) {
            int64_t r = orc_guided_tester(sets, w, len, ed, 1, post4, post_len, bailout, o, list + n, cap - n, n_probes);
            if (r < 0) goto exception;
            if (n + r <= cap) { n += r; break; }
            cap = (n + r) * 2;                      /* list too small: grow and run this window again */
            list = (orc_guided_hit *)realloc(list, (size_t)cap * sizeof(*list));
        }
    }
    out->n_raw = (int32_t)n;
    for (int64_t i = 0; i < n && raw_out && i < raw_cap; i++) raw_out[i] = list[i];
    if (n > 0) {
        gsort_item *it = (gsort_item *)malloc((size_t)n * sizeof(*it));
        for (int64_t i = 0; i < n; i++) {
            it[i].h = list[i]; it[i].idx = i; it[i].bc = bc_flavour;
            if ((list[i].where & ORC_W_GENE) && list[i].n_sub + list[i].n_ins + list[i].n_del < out->min_err_gene)
                out->min_err_gene = list[i].n_sub + list[i].n_ins + list[i].n_del;
        }
        if (n > 1) qsort(it, (size_t)n, sizeof(*it), gsort_cmp);                        /* L52-L56 (size() > 1 only) */
        int nd = 0;
        for (int64_t i = 0; i < n && nd < 2; i++) {                                     /* distinct(): equals = (sequence, seqlength) */
            if (nd == 1 && it[i].h.seq == out->seq[0]) continue;
            out->seq[nd] = it[i].h.seq; out->n_sub[nd] = it[i].h.n_sub; out->n_ins[nd] = it[i].h.n_ins;
            out->n_del[nd] = it[i].h.n_del; out->offset[nd] = it[i].h.offset; out->where[nd] = it[i].h.where;
            nd++;
        }
        out->n_distinct = (uint8_t)nd;
        free(it);
    }
    free(list);
    return;
exception:
    free(list);
    memset(out, 0, sizeof(*out));
    out->min_err_gene = INT32_MAX;
    out->flags = ORC_G_EXCEPTION;
}

void orc_guided_batch(const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups, const uint64_t *all_keys,
                      int64_t n_all, int all_ed, const uint64_t *empty_keys, int64_t n_empty, int empty_ed, int bc_flavour, int len,
                      int plusminus, int bailout, int post_len, const uint8_t *slices, int stride, int slice_len,
                      const int32_t *anchor, const int32_t *group_id, const int32_t *ed, int64_t n, orc_guided_result *out,
                      orc_guided_hit *raw_out, int64_t raw_cap, int64_t *n_probes_total, int n_threads)
{
    orc_set **gs = (orc_set **)calloc((size_t)(n_groups > 0 ? n_groups : 1), sizeof(orc_set *));
    for (int64_t g = 0; g < n_groups; g++) {
        const int64_t m = group_offsets[g + 1] - group_offsets[g];
        if (m > 0) gs[g] = orc_set_new(group_keys + group_offsets[g], m);              /* isEmpty() -> null (L299) */
    }
    orc_set *all = all_keys ? orc_set_new(all_keys, n_all) : NULL;
    orc_set *empty = empty_keys ? orc_set_new(empty_keys, n_empty) : NULL;
    int64_t probes = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : probes)
#endif
    for (int64_t i = 0; i < n; i++) {
        orc_guided_sets s;
        s.group = (group_id[i] >= 0 && group_id[i] < n_groups) ? gs[group_id[i]] : NULL;
        s.all = all; s.all_ed = all_ed; s.empty = empty; s.empty_ed = empty_ed; s.bc_flavour = bc_flavour;
        int64_t p = 0;
        orc_guided_query(&s, bc_flavour, len, ed[i], plusminus, bailout, post_len, slices + (size_t)i * stride, slice_len, anchor[i],
                         &out[i], raw_out ? raw_out + i * raw_cap : NULL, raw_cap, &p);
        probes += p;
    }
    if (n_probes_total) *n_probes_total += probes;
    for (int64_t g = 0; g < n_groups; g++) orc_set_free(gs[g]);
    free(gs);
    orc_set_free(all); orc_set_free(empty);
}
