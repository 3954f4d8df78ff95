// umi_core.cuh — per-pair logic of the UMI distance kernel, __host__ __device__ (shared with tests/host_sim).
//
// Reference: ClusteringEditDistanceBase.calcEditDistances + calcBestEditDistance
//   (F!com/rw/clustering/ClusteringEditDistanceBase.class, ClusteringEditDistanceBase.java:L297-L350, L67-L80),
//   LevenshteinDistance.limitedCompare(threshold 4) (F!com/rw/nanopore/analyzers/apachemod/LevenshteinDistance.class,
//   LevenshteinDistance.java:L220-L283), BestEditDistance packing (…$BestEditDistance.class, L425-L428, L458).
// limitedCompare is a plain Levenshtein distance capped at the threshold (both inputs are umi_len long), so the
// GPU computes the exact distance with Myers/Hyyrö's bit-parallel recurrence (global variant: the horizontal
// delta entering row 0 is +1) on one 32-bit word per comparison and maps d > 4 to 5 like the Java (-1 -> 5).
// The nine window pairs are visited ZERO, PLUSONE, MINUSONE x ZERO, PLUSONE, MINUSONE with strict '<' (EnumSet
// ordinal order, PlusMinusOnePosData.java:L20-L22; calcBestEditDistance L67-L80).
#pragma once
#include "slr_table.cuh"

constexpr int SLR_UMI_MAX_LEN = 14;            // umi_len + 2 codes fit one 64-bit word of nibbles

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

// ---- one thread per read pair (umi_dist.cu), registers only -----------------------------------------------------
// A read = its umi_len+2 codes as 16 bytes (4 little-endian words, code i in byte i).  The ROW read of a pair is also
// given as four bit planes (plane b, bit i = bit b of code i), computed once per read: the Eq mask of a text code t
// is then 4 selects + 3 ANDs, and the Peq tables of the three shifted windows are plain right shifts of it.
SLR_HD uint32_t slr_umi_gather4(uint32_t x)                      // bits 0, 8, 16, 24 of x -> bits 0..3
{
    return ((x & 0x01010101u) * 0x00204081u >> 21) & 15u;
}
SLR_HD void slr_umi_planes(const uint32_t w[4], uint32_t pl[4])
{
#pragma unroll
    for (int b = 0; b < 4; b++)
        pl[b] = slr_umi_gather4(w[0] >> b) | (slr_umi_gather4(w[1] >> b) << 4) | (slr_umi_gather4(w[2] >> b) << 8) |
                (slr_umi_gather4(w[3] >> b) << 12);
}
SLR_HD unsigned long long slr_umi_planes_pack(const uint32_t pl[4])
{
    return (unsigned long long)(pl[0] | (pl[1] << 16)) | ((unsigned long long)(pl[2] | (pl[3] << 16)) << 32);
}

// packed BestEditDistance of the pair (row read given by its planes, column read by its code words); L = umi_len.
// Levenshtein per window pair with Myers/Hyyro's recurrence on one 32-bit word; bits above L-1 hold garbage that only
// flows upwards; the distance is read off the last column: D[L][L] = L + popc(Pv) - popc(Mv) over the low L bits.
template <int L> SLR_HD int32_t slr_umi_best9_planes(unsigned long long row_planes, const uint32_t colw[4])
{
    const uint32_t p0 = (uint32_t)row_planes & 0xFFFFu, p1 = (uint32_t)row_planes >> 16;
    const uint32_t p2 = (uint32_t)(row_planes >> 32) & 0xFFFFu, p3 = (uint32_t)(row_planes >> 48);
    uint32_t eq[L + 2];                                          // eq[c]: row positions whose code equals column code c
#pragma unroll
    for (int c = 0; c < L + 2; c++) {
        const uint32_t t = (colw[c >> 2] >> (8 * (c & 3))) & 15u;
        eq[c] = ((t & 1u) ? p0 : ~p0) & ((t & 2u) ? p1 : ~p1) & ((t & 4u) ? p2 : ~p2) & ((t & 8u) ? p3 : ~p3);
    }
    const uint32_t mask = (1u << L) - 1u;
    int best = 127, b1 = 0, b2 = 0;
#pragma unroll
    for (int x = 0; x < 3; x++) {
        const int i = (x == 0) ? 1 : (x == 1 ? 2 : 0);             // EnumSet order ZERO, PLUSONE, MINUSONE; getValue() = shift + 1
#pragma unroll
        for (int y = 0; y < 3; y++) {
            const int v = (y == 0) ? 1 : (y == 1 ? 2 : 0);
            uint32_t Pv = 0xFFFFFFFFu, Mv = 0;
#pragma unroll
            for (int c = 0; c < L; c++) {
                const uint32_t Eq = eq[v + c] >> i;
                const uint32_t Xv = Eq | Mv;
                const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                const uint32_t Ph = Mv | ~(Xh | Pv);
                const uint32_t Mh = Pv & Xh;
                const uint32_t Ph1 = (Ph << 1) | 1u;               // global alignment: D[0][j] = j
                const uint32_t Mh1 = Mh << 1;
                Pv = Mh1 | ~(Xv | Ph1);
                Mv = Ph1 & Xv;
            }
            int d = L + slr_popc(Pv & mask) - slr_popc(Mv & mask);
            if (d > 4) d = 5;                                        // limitedCompare -> -1 -> 5 (L343)
            if (d < best) { best = d; b1 = i; b2 = v; }
        }
    }
    return slr_umi_pack_best(best, b1, b2);
}
