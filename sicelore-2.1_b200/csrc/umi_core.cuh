// umi_core.cuh — per-pair logic of the UMI distance kernel, __host__ __device__ (shared with tests/host_sim).
//
// Reference: ClusteringEditDistanceBase.calcEditDistances + calcBestEditDistance
//   (F!com/rw/clustering/ClusteringEditDistanceBase.class, ClusteringEditDistanceBase.java:L297-L350, L67-L80),
//   LevenshteinDistance.limitedCompare(threshold 4) (F!com/rw/nanopore/analyzers/apachemod/LevenshteinDistance.class,
//   LevenshteinDistance.java:L220-L283), BestEditDistance packing (…$BestEditDistance.class, L425-L428, L458).
// limitedCompare is a plain Levenshtein distance capped at the threshold (both inputs are umi_len long), so the
// GPU computes the exact distance with Myers/Hyyrö's bit-parallel recurrence (global variant: the horizontal
// delta entering row 0 is +1) on one 32-bit word per comparison and maps d > 4 to 5 like the Java (-1 -> 5).
#pragma once
#include "slr_table.cuh"

constexpr int SLR_UMI_MAX_LEN = 14;            // umi_len + 2 codes fit one 64-bit word of nibbles

// pack umi_len+2 4-bit codes (NucleicAcidByteCodeBase codes, java:L45-L78) into nibbles, code i at bits 4i
SLR_HD unsigned long long slr_umi_pack(const uint8_t *codes, int n)
{
    unsigned long long r = 0;
    for (int i = 0; i < n; i++) r |= (unsigned long long)(codes[i] & 15u) << (4 * i);
    return r;
}

// Peq tables of the three shifted windows (-1, 0, +1) of a row read: peq[s*16 + code] = bitmask of the
// positions i < umi_len with window_s[i] == code
SLR_HD uint32_t slr_umi_peq_entry(unsigned long long packed, int umi_len, int s, uint32_t code)
{
    uint32_t m = 0;
    for (int i = 0; i < umi_len; i++)
        m |= (uint32_t)(((packed >> (4 * (i + s))) & 15ull) == code) << i;
    return m;
}

// exact Levenshtein distance between the pattern described by peq (length m) and the text = nibbles
// t0..t0+m-1 of `text`
SLR_HD int slr_umi_myers(const uint32_t *peq, int m, unsigned long long text, int t0)
{
    const uint32_t mask = slr_lowmask(m), top = 1u << (m - 1);
    uint32_t Pv = mask, Mv = 0;
    int score = m;
    for (int c = 0; c < m; c++) {
        const uint32_t Eq = peq[(text >> (4 * (t0 + c))) & 15ull];
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = ((((Eq & Pv) + Pv) ^ Pv) | Eq);
        uint32_t Ph = Mv | ~(Xh | Pv);
        uint32_t Mh = Pv & Xh;
        score += (Ph & top) ? 1 : 0;
        score -= (Mh & top) ? 1 : 0;
        Ph = (Ph << 1) | 1u;                    // global alignment: D[0][j] = j
        Mh = Mh << 1;
        Pv = (Mh | ~(Xv | Ph)) & mask;
        Mv = Ph & Xv & mask;
    }
    return score;
}

// BestEditDistance(ed, pos1, pos2): value MINUSONE=0 ZERO=1 PLUSONE=2 (java:L425-L428)
SLR_HD int32_t slr_umi_pack_best(int ed, int v1, int v2)
{
    return (int32_t)((uint32_t)(ed & 0xFFFFFF) | (0x08000000u << v1) | (0x01000000u << v2));
}
SLR_HD int32_t slr_umi_transpose(int32_t packed)                 // getTransposedCopy (java:L458)
{
    const uint32_t u = (uint32_t)packed;
    return (int32_t)((u & 0xFFFFFFu) | (((u >> 24) & 7u) << 27) | (((u >> 27) & 7u) << 24));
}
SLR_HD int32_t slr_umi_equality() { return slr_umi_pack_best(0, 1, 1); }   // EQUALITYMATRIX (java:L90-L92)

// best of the 3 x 3 shifted comparisons, visited ZERO, PLUSONE, MINUSONE x ZERO, PLUSONE, MINUSONE with strict '<'
// (EnumSet ordinal order, PlusMinusOnePosData.java:L20-L22; calcBestEditDistance L67-L80)
SLR_HD int32_t slr_umi_best9(const uint32_t *peq_row /* [3][16] */, int umi_len, unsigned long long col_text)
{
    int best = 127, b1 = 0, b2 = 0;
#pragma unroll
    for (int x = 0; x < 3; x++) {
        const int i = (x == 0) ? 1 : (x == 1 ? 2 : 0);             // getValue(): window shift i-1
#pragma unroll
        for (int y = 0; y < 3; y++) {
            const int v = (y == 0) ? 1 : (y == 1 ? 2 : 0);
            int d = slr_umi_myers(peq_row + 16 * i, umi_len, col_text, v);
            if (d > 4) d = 5;                                        // limitedCompare -> -1 -> 5 (L343)
            if (d < best) { best = d; b1 = i; b2 = v; }
        }
    }
    return slr_umi_pack_best(best, b1, b2);
}
