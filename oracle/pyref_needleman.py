"""TEST INFRASTRUCTURE — Python restatement of the alignment step that decides MORE_THAN_ONE_MATCH in the Illumina-guided search:
IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI (F!com/rw/umifinder/analyzers/IlluminaBarcodeUMIAnalyzerBase.class, …java:L66-L86) aligns the
best and the second-best candidate to the read window they were found in (NeedlemanWunsch, T!com/rw/nuc/alignment/needleman/{DynamicProgramming,
SequenceAlignment,NeedlemanWunsch}.class) and compares NeedlemanMatch.countNeedlemanErrorsInRead (F!com/rw/nanopore/analyzers/NeedlemanMatch.class,
NeedlemanMatch.java:L68-L86) of the two alignments; a difference of 0 sets the flag, which turns `found` off (IlluminaUMIanalyzer.java:L203-L220).
The product's implementation is csrc/slr_needleman.cpp; pinned by oracle/make_ref_needleman.py -> tests/golden/ref_needleman.npz."""

DEFAULT_SCORES = dict(leading_gap_1=-4, leading_gap_2=-5, trailing_gap_1=-5, trailing_gap_2=-5, indel=-5, mismatch=-5, match=5)      # NeedlemanScores.java:L44-L56


def align(template, read, scores=None):
    """NeedlemanWunsch(template, read, scores).getAlignmentString(): (template row, pattern, read row); '-' = gap, 'x' = mismatch / gap, '.' = match.
    Columns = template (sequence1), rows = read (sequence2).  Ties: diagonal first, then the cell above unless the left one scores strictly
    higher (NeedlemanWunsch.java:L55-L80); the traceback starts in the last cell (global alignment; the trailing-gap scores are not used)."""
    s = dict(DEFAULT_SCORES, **(scores or {}))
    T, R = len(template), len(read)
    sc = [[0] * (T + 1) for _ in range(R + 1)]
    pv = [[None] * (T + 1) for _ in range(R + 1)]
    for c in range(1, T + 1):
        sc[0][c], pv[0][c] = c * s["leading_gap_2"], (0, c - 1)                # NeedlemanWunsch.java:L106-L122
    for r in range(1, R + 1):
        sc[r][0], pv[r][0] = r * s["leading_gap_1"], (r - 1, 0)
    for r in range(1, R + 1):
        for c in range(1, T + 1):
            row_space = sc[r - 1][c] + s["indel"]
            col_space = sc[r][c - 1] + s["indel"]
            diag = sc[r - 1][c - 1] + (s["match"] if read[r - 1] == template[c - 1] else s["mismatch"])
            if row_space >= col_space:
                sc[r][c], pv[r][c] = (diag, (r - 1, c - 1)) if diag >= row_space else (row_space, (r - 1, c))
            else:
                sc[r][c], pv[r][c] = (diag, (r - 1, c - 1)) if diag >= col_space else (col_space, (r, c - 1))
    a1, a2 = [], []
    cur = (R, T)
    while pv[cur[0]][cur[1]] is not None:                                     # SequenceAlignment.getTraceback (…java:L108-L120)
        p = pv[cur[0]][cur[1]]
        a2.insert(0, read[cur[0] - 1] if cur[0] - p[0] == 1 else "-")
        a1.insert(0, template[cur[1] - 1] if cur[1] - p[1] == 1 else "-")
        cur = p
    pat = "".join("x" if x == "-" or y == "-" or x != y else "." for x, y in zip(a1, a2))
    return "".join(a1), pat, "".join(a2)


def count_errors(match, pattern, read):
    """NeedlemanMatch.countNeedlemanErrorsInRead: (insertions, deletions, substitutions, total); gaps at the END of the read row are not counted"""
    ins = dele = sub = 0
    for i in range(len(match)):
        if pattern[i] == "x":
            if match[i] == "-":
                ins += 1
            elif read[i] == "-":
                dele += 1
            else:
                sub += 1
    i = len(read)
    while read[i - 1] == "-":                                                 # StringIndexOutOfBounds on an all-gap row: cannot happen for len > 0
        i -= 1
    dele -= len(read) - i
    return ins, dele, sub, ins + dele + sub


def n_errors(template, read, scores=None):
    return count_errors(*align(template, read, scores))[3]
