// slr_needleman.cpp — host finishing step of the Illumina-guided search (seam S4): the alignment comparison that decides MORE_THAN_ONE_MATCH.
//
// The guided kernel delivers, per read, the first two entries of the sorted, distinct match list.  The reference then aligns each of the two to
// the read window it was found in and compares the error counts of the two alignments; equal counts flag the read MORE_THAN_ONE_MATCH, which
// turns `found` off (IlluminaUMIanalyzer.java:L203-L220).  Two 12- or 16-mers per read — host work, like in the reference.  Restated from:
//   IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI   F!com/rw/umifinder/analyzers/IlluminaBarcodeUMIAnalyzerBase.class (…java:L66-L86)
//   NeedlemanWunsch / SequenceAlignment / DynamicProgramming  T!com/rw/nuc/alignment/needleman/*.class (NeedlemanWunsch.java:L55-L122,
//                                                             SequenceAlignment.java:L102-L151, DynamicProgramming.java:L57-L98)
//   NeedlemanMatch.countNeedlemanErrorsInRead                 F!com/rw/nanopore/analyzers/NeedlemanMatch.class (NeedlemanMatch.java:L68-L86)
// Pinned by tests/golden/ref_needleman.npz (600 alignments run from the class files) and the nMismatchDiffBestvsSecondBest of the findUMI vectors.
#include <cstdint>
#include <cstring>

#include "../../include/sicelore_host.h"

extern "C" int slr_multi_fail(int code, const char *msg);      // slr_api.cu: sets the thread-local error message

namespace {

const slr_needleman_scores DEFAULT_SCORES = {-4, -5, -5, -5, -5, -5, 5};      // NeedlemanScores.java:L44-L56

constexpr int MAXL = 32;

inline int base_at(uint64_t v, int len, int i) { return (int)((v >> (2 * (len - 1 - i))) & 3); }

// counts[0..3] = insertions, deletions, substitutions, total of NeedlemanMatch for NeedlemanWunsch(template, read, scores)
void nw_errors(uint64_t templ, uint64_t read, int len, const slr_needleman_scores &s, int32_t counts[4])
{
    int sc[MAXL + 1][MAXL + 1];
    uint8_t pv[MAXL + 1][MAXL + 1];                            // 0 = none, 1 = diagonal, 2 = above (row - 1), 3 = left (col - 1)
    const int T = len, R = len;                                // columns = template (sequence1), rows = read (sequence2)
    sc[0][0] = 0; pv[0][0] = 0;
    for (int c = 1; c <= T; c++) { sc[0][c] = c * s.leading_gap_2; pv[0][c] = 3; }      // NeedlemanWunsch.java:L106-L122
    for (int r = 1; r <= R; r++) { sc[r][0] = r * s.leading_gap_1; pv[r][0] = 2; }
    for (int r = 1; r <= R; r++)
        for (int c = 1; c <= T; c++) {                         // fillInCell (L55-L80): diagonal first, then above unless left is strictly higher
            int row_space = sc[r - 1][c] + s.indel, col_space = sc[r][c - 1] + s.indel;
            int diag = sc[r - 1][c - 1] + (base_at(read, len, r - 1) == base_at(templ, len, c - 1) ? s.match : s.mismatch);
            if (row_space >= col_space) {
                if (diag >= row_space) { sc[r][c] = diag; pv[r][c] = 1; } else { sc[r][c] = row_space; pv[r][c] = 2; }
            } else {
                if (diag >= col_space) { sc[r][c] = diag; pv[r][c] = 1; } else { sc[r][c] = col_space; pv[r][c] = 3; }
            }
        }
    // traceback from the last cell (SequenceAlignment.java:L108-L120), columns collected back to front
    int8_t a1[2 * MAXL + 2], a2[2 * MAXL + 2];                 // base 0..3 or -1 = gap
    int n = 0, r = R, c = T;
    while (pv[r][c]) {
        int p = pv[r][c];
        a2[n] = (int8_t)(p != 3 ? base_at(read, len, r - 1) : -1);
        a1[n] = (int8_t)(p != 2 ? base_at(templ, len, c - 1) : -1);
        n++;
        if (p != 3) r--;
        if (p != 2) c--;
    }
    int ins = 0, del = 0, sub = 0;                             // NeedlemanMatch.java:L68-L86 (index n - 1 is the first column)
    for (int i = 0; i < n; i++) {
        if (a1[i] < 0) ins++;
        else if (a2[i] < 0) del++;
        else if (a1[i] != a2[i]) sub++;
    }
    for (int i = 0; i < n && a2[i] < 0; i++) del--;            // gaps at the END of the read row are not counted
    counts[0] = ins; counts[1] = del; counts[2] = sub; counts[3] = ins + del + sub;
}

inline int code_of(uint8_t ch)                                 // A G C T -> 0 1 2 3 (the reference's 2-bit code), anything else -1
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'G': case 'g': return 1;
    case 'C': case 'c': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

}   // namespace

extern "C" {

int slr_needleman_errors(uint64_t template2bit, uint64_t read2bit, int len, const slr_needleman_scores *scores, int32_t *counts_out)
{
    if (!counts_out) return slr_multi_fail(SLR_E_INVALID, "slr_needleman_errors: counts_out is NULL");
    if (len < 1 || len > MAXL) return slr_multi_fail(SLR_E_UNSUPPORTED, "slr_needleman_errors: 1 <= len <= 32");
    nw_errors(template2bit, read2bit, len, scores ? *scores : DEFAULT_SCORES, counts_out);
    return SLR_OK;
}

int slr_guided_mismatch_diff(const slr_guided_result *res, int64_t n, const uint8_t *slices, int stride, int slice_len, const int32_t *anchor,
                             int seq_len, const slr_needleman_scores *scores, int32_t *diff_out)
{
    if (n < 0 || (n > 0 && (!res || !slices || !anchor || !diff_out))) return slr_multi_fail(SLR_E_INVALID, "slr_guided_mismatch_diff: NULL argument / n < 0");
    if (seq_len < 1 || seq_len > MAXL || slice_len > stride || slice_len < seq_len) return slr_multi_fail(SLR_E_INVALID, "slr_guided_mismatch_diff: bad seq_len / slice_len / stride");
    const slr_needleman_scores &s = scores ? *scores : DEFAULT_SCORES;
    for (int64_t i = 0; i < n; i++) {
        diff_out[i] = SLR_G_NO_SECOND;
        if (res[i].flags || res[i].n_distinct < 2) continue;
        int32_t err[2];
        bool ok = true;
        for (int k = 0; k < 2 && ok; k++) {
            int start = anchor[i] + res[i].offset[k];          // the window the entry's tester started from = its unMutatedSeq
            if (start < 0 || start + seq_len > slice_len) { ok = false; break; }
            uint64_t w = 0;
            for (int j = 0; j < seq_len; j++) {
                int cde = code_of(slices[i * (int64_t)stride + start + j]);
                if (cde < 0) { ok = false; break; }
                w = (w << 2) | (uint64_t)cde;
            }
            if (!ok) break;
            int32_t cnt[4];
            nw_errors(res[i].seq[k], w, seq_len, s, cnt);
            err[k] = cnt[3];
        }
        if (!ok) return slr_multi_fail(SLR_E_INVALID, "slr_guided_mismatch_diff: a record's window lies outside its slice or holds a non-ACGT base "
                                                      "(not the slices / anchors the records were computed from?)");
        diff_out[i] = err[1] - err[0];                         // nMismatchDiffBestvsSecondBest (IlluminaBarcodeUMIAnalyzerBase.java:L79)
    }
    return SLR_OK;
}

}   // extern "C"
