// slr_table.cuh — device-resident barcode search set ("digit-group bucket tables").
//
// What it replaces: java.util.Set<Long>.contains on the fastutil key set of BarcodesMapForBCfinding,
// probed once per mutant by BarcodeMatchTester.checkMatchWithTestSets
// (F!com/rw/nanoporereadscanner/analyzers/BarcodeMatchTester.class, BarcodeMatchTester.java:L367-L374).
//
// Layout (B200: everything is sized to stay L2-resident, 4 x 16 MB for a 3 M list):
//   a 16-nt barcode is 16 base-4 digits d0..d15 (d0 most significant, A=0 G=1 C=2 T=3).  Table g (g=0..3)
//   "ignores" digit group g = digits 4g..4g+3: a key is split into  pat = those 8 bits  and  rest = the other
//   24 bits.  rest goes through a 24-bit bijection; its top `bbits` bits select a 32-byte bucket (one L2
//   sector), the remaining 24-bbits bits are a tag.  A bucket holds 16 slots: bytes 0..15 = 0x80 | tag of
//   slot i (0 = empty), bytes 16..31 = pat of slot i   (exact: (bucket, tag, pat) <-> key is a bijection).
//   => all single-edit neighbours of a node whose edit falls into digit group g share ONE bucket of table g,
//   so one 32-byte load tests up to 16 mutants of the reference's enumeration at once.
//   Buckets that overflow spill into a small sorted stash (checked only when a bucket is completely full).
//
// Everything here is __host__ __device__ so that tests/host_sim can run the same logic on the CPU.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

struct SlrTableDev {
    const uint4 *bk;               // the four tables back to back: table g starts at bucket g << bbits; 2 x uint4 per bucket
    const uint32_t *st_bucket;     // stash of all four tables: sorted (g << 24 | bucket) ids
    const uint16_t *st_slot;       //        and their slots
    int st_total;                  // stash entries (0 for almost every list: the stash code is then skipped)
    int bbits;                     // log2(#buckets per table), 17..22  (tag bits = 24 - bbits <= 7)
    const uint32_t *ix_keys;       // key -> index map (open addressing, linear probing)
    const int32_t *ix_vals;        //   -1 = empty
    uint32_t ix_mask;
    const int32_t *rank;           // CountsRank.rank per index (may be null)
    unsigned long long *counts;    // n x 3 assigned-read counters by ED (BarcodeCounts.addCountForEd)
    long long n;
};

#define SLR_HD __host__ __device__ __forceinline__

// ---- intrinsics with host equivalents ----------------------------------------------------------------
SLR_HD int slr_ffs(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return x ? __builtin_ctz(x) + 1 : 0;
#endif
}
SLR_HD int slr_popc(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
SLR_HD int slr_clz(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
SLR_HD uint32_t slr_brev(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
SLR_HD uint32_t slr_byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, s);
#else
    const uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
template <typename T> SLR_HD T slr_ldg(const T *p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

SLR_HD uint32_t slr_lowmask(int nbits) { return nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u); }

// 24-bit bijection (odd multipliers mod 2^24, xorshifts)
SLR_HD uint32_t slr_mix24(uint32_t id)
{
    uint32_t m = (id * 0x9E3779u) & 0xFFFFFFu;
    m ^= m >> 12;
    m = (m * 0x85EBCBu) & 0xFFFFFFu;
    m ^= m >> 11;
    return m;
}

// split a 32-bit key for table g
SLR_HD uint32_t slr_key_pat(uint32_t key, int g) { return (key >> (24 - 8 * g)) & 0xFFu; }
SLR_HD uint32_t slr_key_rest(uint32_t key, int g)
{
    const int lo_bits = 24 - 8 * g;
    const uint32_t hi = g == 0 ? 0u : (key >> (32 - 8 * g));
    return (hi << lo_bits) | (key & slr_lowmask(lo_bits));
}
SLR_HD uint32_t slr_key_join(uint32_t rest, uint32_t pat, int g)
{
    const int lo_bits = 24 - 8 * g;
    const uint32_t hi = rest >> lo_bits;
    return (g == 0 ? 0u : (hi << (32 - 8 * g))) | (pat << lo_bits) | (rest & slr_lowmask(lo_bits));
}
SLR_HD uint32_t slr_ix_hash(uint32_t key) { return (key * 0x9E3779B1u) ^ (key >> 15); }

// One bucket = 16 slots in two uint4: a = the 16 tag bytes (0x80 | tag, 0 = empty slot), b = the 16 pattern bytes.
// (Tags and patterns sit in separate halves so that the tag filter is four word-wide byte compares.)
struct SlrBucket {
    uint4 a, b;
};

// (no dynamically indexed member in SlrTableDev: a kernel parameter indexed at run time is copied to local memory)
SLR_HD SlrBucket slr_load_bucket(const SlrTableDev &t, int g, uint32_t bucket)
{
    const uint4 *p = t.bk + 2 * (((size_t)g << t.bbits) + bucket);
    SlrBucket r;
#ifdef __CUDA_ARCH__
    // one 256-bit read-only load = one 32-byte L2 sector per lane (sm_100: LDG.E.256.CONSTANT).  -DSLR_BC_L2HINT=1 keeps the table's
    // sectors with L2::evict_last and streams the read slices / records with evict-first (.cs) accesses; measured on B200 (10 M reads,
    // 3 M list, ED 2): 48.33 ms with the hints, 48.06 ms without, L2 sector hit rate 76 % either way (profiles/r2_bc_variants.txt) — off.
#if defined(SLR_BC_L2HINT) && SLR_BC_L2HINT == 1
    asm("ld.global.nc.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
        : "l"(p));
#else
    r.a = p[0];
    r.b = p[1];
#endif
    return r;
}

// bit 7 of every byte of the result set <=> that byte of x equals the byte replicated in `want` (exact, carry-free)
SLR_HD uint32_t slr_eq_bytes(uint32_t x, uint32_t want)
{
    const uint32_t h = x ^ want;
    return ~(((h & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | h) & 0x80808080u;
}
// bit 7 of every byte of the result set <=> the high OR the low nibble of that byte of h is zero (exact, carry-free per nibble)
#ifndef SLR_NIBBLE_FILTER
#define SLR_NIBBLE_FILTER 1
#endif
SLR_HD uint32_t slr_nibble_ok(uint32_t h)
{
    const uint32_t z = ~(((h & 0x77777777u) + 0x77777777u) | h) & 0x88888888u;
    return z | (z << 4);
}
// Bit 8*byte + word of the result set  <=>  slot 4*word + byte (byte `byte` of tag word `word`) is valid and
// carries `tag` (the four per-word byte masks are interleaved with four shifts; the order of the set bits is
// irrelevant to the callers, which take a minimum over all matches).
SLR_HD uint32_t slr_tag_match(const SlrBucket &k, uint32_t tag)
{
    const uint32_t want = (0x80u | tag) * 0x01010101u;
    return (slr_eq_bytes(k.a.x, want) >> 7) | (slr_eq_bytes(k.a.y, want) >> 6) | (slr_eq_bytes(k.a.z, want) >> 5) |
           (slr_eq_bytes(k.a.w, want) >> 4);
}

// pattern byte of the slot named by bit b of slr_tag_match
SLR_HD uint32_t slr_bucket_pat(const SlrBucket &k, int b)
{
    const uint32_t w = (b & 2) ? ((b & 1) ? k.b.w : k.b.z) : ((b & 1) ? k.b.y : k.b.x);
    return (w >> (b & 24)) & 0xFFu;
}

SLR_HD bool slr_bucket_full(const SlrBucket &k) { return ((k.a.x & k.a.y & k.a.z & k.a.w) & 0x80808080u) == 0x80808080u; }

// first stash entry of bucket id `gb` = g << 24 | bucket (lower bound); its entries end where st_bucket != gb
SLR_HD int slr_stash_lower(const uint32_t *st_bucket, int st_total, uint32_t gb)
{
    int lo = 0, hi = st_total;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (slr_ldg(st_bucket + mid) < gb) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exact membership of one key (the ED-0 probe, BarcodeMatchTester.java:L204)
SLR_HD bool slr_contains_in(const SlrTableDev &t, const SlrBucket &k, uint32_t bucket, uint32_t tag, uint32_t pat)
{
    const uint32_t pw = pat * 0x01010101u;
    const uint32_t want = (0x80u | tag) * 0x01010101u;
    uint32_t hit = slr_eq_bytes(k.a.x, want) & slr_eq_bytes(k.b.x, pw);
    hit |= slr_eq_bytes(k.a.y, want) & slr_eq_bytes(k.b.y, pw);
    hit |= slr_eq_bytes(k.a.z, want) & slr_eq_bytes(k.b.z, pw);
    hit |= slr_eq_bytes(k.a.w, want) & slr_eq_bytes(k.b.w, pw);
    if (hit) return true;
    if (t.st_total > 0 && slr_bucket_full(k)) {
        const uint32_t sw = 0x8000u | (tag << 8) | pat;
        for (int i = slr_stash_lower(t.st_bucket, t.st_total, bucket); i < t.st_total && slr_ldg(t.st_bucket + i) == bucket; i++)
            if ((uint32_t)slr_ldg(t.st_slot + i) == sw) return true;
    }
    return false;
}
SLR_HD bool slr_contains(const SlrTableDev &t, uint32_t key)
{
    const uint32_t m = slr_mix24(slr_key_rest(key, 0));
    const int tb = 24 - t.bbits;
    const uint32_t bucket = m >> tb, tag = m & ((1u << tb) - 1u);
    return slr_contains_in(t, slr_load_bucket(t, 0, bucket), bucket, tag, slr_key_pat(key, 0));
}

// key -> index in the caller's barcode array (for rank / counters); -1 if absent
SLR_HD int slr_index_of(const SlrTableDev &t, uint32_t key)
{
    uint32_t h = slr_ix_hash(key) & t.ix_mask;
    while (true) {
        const int v = slr_ldg(t.ix_vals + h);
        if (v < 0) return -1;
        if (slr_ldg(t.ix_keys + h) == key) return v;
        h = (h + 1) & t.ix_mask;
    }
}
