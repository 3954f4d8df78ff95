"""Golden vectors of the alignment step behind MORE_THAN_ONE_MATCH FROM THE REFERENCE'S OWN CLASS FILES: new NeedlemanWunsch(new
NucleicAcidOneBytePerBase(candidate), new NucleicAcidOneBytePerBase(read window), scores).getAlignmentString() and new NeedlemanMatch(match,
pattern, read) (what Match.addNeedlemanAlignment builds; IlluminaBarcodeUMIAnalyzerBase.java:L66-L79) run by oracle/minijvm.py.
Frozen in tests/golden/ref_needleman.npz.

    python oracle/make_ref_needleman.py [n_pairs]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import minijvm as J  # noqa: E402
from oracle import make_ref_vectors as M  # noqa: E402
from oracle import pyref_needleman as P  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_needleman.npz")
OB = "com/rw/nuc/encoding/onebyte/NucleicAcidOneBytePerBase"
NW = "com/rw/nuc/alignment/needleman/NeedlemanWunsch"
NM = "com/rw/nanopore/analyzers/NeedlemanMatch"
SC = "com/rw/nuc/alignment/needleman/NeedlemanScores"


def run_pair(vm, a, b, L, scores):
    ta = vm.construct(M.T2, "(JI)V", J.L(M.pack(a)), L)
    tb = vm.construct(M.T2, "(JI)V", J.L(M.pack(b)), L)
    oa = vm.construct(OB, "(L%s;)V" % M.T2, ta)
    ob = vm.construct(OB, "(L%s;)V" % M.T2, tb)
    if scores is None:
        s = vm.construct(SC, "()V")
    else:
        s = vm.construct(SC, "(IIIIIII)V", *scores)
    nw = vm.construct(NW, "(L%s;L%s;L%s;)V" % (M.ONEBYTE, M.ONEBYTE, SC), oa, ob, s)
    al = vm.call_virtual(nw, "getAlignmentString", "()[Ljava/lang/String;")
    m = vm.construct(NM, "(Ljava/lang/String;Ljava/lang/String;Ljava/lang/String;)V", al.a[0], al.a[1], al.a[2])
    return al.a[0], al.a[1], al.a[2], m.f["insertionsNeedleman"], m.f["deletionsNeedleman"], m.f["substitutionsNeedleman"], int(m.f["nMismatchesInAlignment"])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    vm = J.VM(M.JARS + [M.REF + "/lib/commons-lang3-3.17.0.jar"])
    rng = np.random.default_rng(20261020)
    t0 = time.time()
    rows, bad = [], 0
    for t in range(n):
        L = [12, 16, 12, 16, 10, 8][t % 6]
        a = M.rseq(rng, L, "AAAGCT" if t % 7 == 0 else "AGCT")
        k = int(rng.integers(0, 5))
        b = a
        for _ in range(k):                                                    # substitutions, insertions and deletions at fixed length, like the engine's mutants
            b = (M.mutate(rng, b, 1) + M.rseq(rng, 2))[:L]
        if t % 11 == 0:
            b = a[1:] + M.rseq(rng, 1)                                        # a pure shift: leading / trailing gaps
        if t % 13 == 0:
            b = M.rseq(rng, L)
        custom = t % 5 == 4
        sc = None if not custom else [int(rng.integers(-6, 0)), int(rng.integers(-6, 0)), -5, -5, int(rng.integers(-7, -1)), int(rng.integers(-7, -1)), int(rng.integers(1, 7))]
        m, p, r, ins, dele, sub, tot = run_pair(vm, a, b, L, sc)
        rows.append((a, b, L, m, p, r, ins, dele, sub, tot, sc or [0] * 7, int(custom)))
        pd = None if sc is None else dict(zip(("leading_gap_1", "leading_gap_2", "trailing_gap_1", "trailing_gap_2", "indel", "mismatch", "match"), sc))
        exp = P.align(a, b, pd)
        if exp != (m, p, r) or P.count_errors(*exp) != (ins, dele, sub, tot):
            bad += 1
            print("    PYREF DIFFERS", t, (a, b), exp, (m, p, r), P.count_errors(*exp), (ins, dele, sub, tot))
    np.savez_compressed(OUT, template=np.array([r[0] for r in rows]), read=np.array([r[1] for r in rows]), L=np.array([r[2] for r in rows], dtype=np.int32),
                        match=np.array([r[3] for r in rows]), pattern=np.array([r[4] for r in rows]), read_row=np.array([r[5] for r in rows]),
                        counts=np.array([r[6:10] for r in rows], dtype=np.int32), scores=np.array([r[10] for r in rows], dtype=np.int32),
                        custom=np.array([r[11] for r in rows], dtype=np.uint8))
    print("NeedlemanWunsch + NeedlemanMatch: %d pairs, error histogram %s, pyref differs on %d, %.0f s, %d bytecodes" %
          (len(rows), np.bincount([r[9] for r in rows]).tolist(), bad, time.time() - t0, vm.n_insn))


if __name__ == "__main__":
    main()
