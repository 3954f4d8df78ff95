"""Golden vectors produced BY THE REFERENCE ITSELF: the class files of /root/reference/Jar are executed by oracle/minijvm.py (there is no
JDK in the image) and the inputs / outputs are frozen in tests/golden/ref_*.npz.  Run in the build container only:

    python oracle/make_ref_vectors.py

What runs is the reference's bytecode, unmodified: NucleicAcidTwoBitPerBase (pack, reverse complement, the three mutation primitives),
BarcodeMatchTester.doJob (second-pass and collision-tester settings), UMInuc/BCnucTwoBitPerBaseEDtester.matchesSeqEditDistance
(Illumina-guided engine incl. bailout) and LevenshteinDistance.apply (thresholded UMI distance).  Containers that are not in the two
jars (java.util, eclipse-collections, fastutil, the Illumina data holders) are Python shims with membership semantics (minijvm.py).
tests/test_ref_vectors.py checks the C oracle (and the GPU path, through the oracle) against these vectors."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import minijvm as J  # noqa: E402

REF = "/root/reference/Jar"
JARS = [REF + "/NanoporeBC_UMI_finder-2.1.jar", REF + "/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar"]
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
T2 = "com/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase"
ONEBYTE = "com/rw/nuc/encoding/onebyte/NucleicAcidInmutableOneBytePerBase"
BMT = "com/rw/nanoporereadscanner/analyzers/BarcodeMatchTester"
UMIT = "com/rw/nuc/encoding/TwoBit/ed/UMInucTwoBitPerBaseEDtester"
BCT = "com/rw/nuc/encoding/TwoBit/ed/BCnucTwoBitPerBaseEDtester"
LEV = "com/rw/nanopore/analyzers/apachemod/LevenshteinDistance"
M64 = (1 << 64) - 1
BASES = "AGCT"


def carr(s):
    a = J.JArr("C", len(s), 0)
    a.a = [ord(c) for c in s]
    return a


def barr(v):
    a = J.JArr("B", len(v), 0)
    a.a = [int(x) - 256 if int(x) > 127 else int(x) for x in v]
    return a


def rseq(rng, n, alpha="AGCT"):
    return "".join(alpha[i] for i in rng.integers(0, len(alpha), n))


def primitives(vm, rng):
    rows, seqs = [], []
    for L in (12, 16):
        for t in range(60):
            s = rseq(rng, L, "AAAAGCT" if t % 3 == 0 else "AGCT")
            if t % 10 == 9:
                p = int(rng.integers(L))
                s = s[:p] + "N" + s[p + 1:]
            h = vm.call_static(T2, "getLongHashForSeq", "([C)J", carr(s))
            rc = vm.call_virtual(vm.construct(T2, "(JI)V", J.L(h), L), "reverseComplement", "()Lcom/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase;").f["sequence"]
            seqs.append((s, L, int(h) & M64, int(rc) & M64))
            out = J.JArr("J", 4, J.L(0))
            for p in range(L):
                vm.call_static(T2, "getLongHashReplaceByteDeg", "(J[JII)V", J.L(h), out, p, L)
                rep = [int(x) & M64 for x in out.a]
                if p < L - 1:
                    vm.call_static(T2, "getLongHashInsertByteDeg", "(J[JII)V", J.L(h), out, p, L)
                    ins = [int(x) & M64 for x in out.a]
                else:
                    ins = [0, 0, 0, 0]
                dl = [int(vm.call_static(T2, "getLongHashdeleteByte", "(JBII)J", J.L(h), c4, p, L)) & M64 for c4 in (1, 2, 4, 8, 15)]
                rows.append([L, int(h) & M64, p, int(rc) & M64] + rep + ins + dl)
    return np.array(rows, dtype=np.uint64), seqs


def onebyte(vm, s):
    return vm.construct(ONEBYTE, "(Ljava/lang/CharSequence;)V", s)


def run_dojob(vm, keys, w, L, ed, skip_full, post, do_next, offset):
    """new BarcodeMatchTester(ed, skipFullMatches, allowIndels = true, searchSet, offset, L, postSeq, doNext).doJob(seq)"""
    t = vm.construct(BMT, "(IZZLjava/util/Set;SILcom/rw/nuc/encoding/onebyte/NucleicAcidInmutableOneBytePerBase;Z)V", ed, int(skip_full), 1,
                     J.PySet(keys), offset, L, None if post is None else onebyte(vm, post), int(do_next))
    seq = vm.construct(T2, "(JI)V", J.L(w), L)
    m = vm.call_virtual(t, "doJob", "(Lcom/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase;)Lcom/rw/nanoporereadscanner/analyzers/BarcodeMatchTester$Matches;", seq)
    out = []
    if m is not None:
        for o in m.native.items():
            f = o.f
            out.append((int(f["readSeq"]) & M64, int(f["matchingBC"]) & M64, f["editDistance"], f["substitutions"], f["insertions"], f["deletions"],
                        f["offsetFromPredicted"]))
    return out


def pack(s):
    v = 0
    for ch in s:
        v = (v << 2) | BASES.index(ch)
    return v


def unpack(v, L):
    return "".join(BASES[(v >> (2 * (L - 1 - i))) & 3] for i in range(L))


def mutate(rng, s, k):
    s = list(s)
    for _ in range(k):
        op, p = int(rng.integers(3)), int(rng.integers(len(s)))
        if op == 0:
            s[p] = BASES[int(rng.integers(4))]
        elif op == 1:
            s.insert(p, BASES[int(rng.integers(4))])
        else:
            del s[p]
    return "".join(s)


def dojob_cases(vm, rng, n_cases, L=16):
    """windows planted 0..3 edits from list members (so that every ED level and the first-hit-wins order are exercised)"""
    cases = []
    for t in range(n_cases):
        alpha = "AAAAGCT" if t % 4 == 0 else "AGCT"
        keys = sorted({pack(rseq(rng, L, alpha)) for _ in range(40)})
        base = unpack(keys[int(rng.integers(len(keys)))], L)
        wstr = (mutate(rng, base, int(rng.integers(0, 4))) + rseq(rng, 8))[:L]
        # plant extra neighbours of the window itself: competing hits at ED 1 / 2
        for _ in range(int(rng.integers(0, 6))):
            keys.append(pack((mutate(rng, wstr, int(rng.integers(1, 3))) + rseq(rng, 4))[:L]))
        keys = sorted(set(keys))
        ed = 1 if t % 5 == 0 else 2
        mode = t % 3                      # 0/1: second pass (post given, doNext), 2: collision tester (no post, !doNext, skipFullMatches)
        post = None if mode == 2 else rseq(rng, 5, "AGCTN" if t % 7 == 0 else "AGCT")
        off = int(rng.integers(-2, 3))
        if mode == 2:
            keys = sorted(set(keys) | {pack(wstr)})
        res = run_dojob(vm, keys, pack(wstr), L, ed, mode == 2, post, mode != 2, off)
        cases.append(dict(keys=keys, w=pack(wstr), ed=ed, mode=mode, post=post or "", off=off, res=res))
    return cases


def guided_cases(vm, rng, n_cases):
    """UMInucTwoBitPerBaseEDtester / BCnucTwoBitPerBaseEDtester.matchesSeqEditDistance: the whole ordered ArrayList"""
    params_cls = vm.load("com/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams")
    cases = []
    for t in range(n_cases):
        bc = t % 3 == 2
        L = 16 if bc else 12
        ed = [1, 2, 2, 3][t % 4] if not bc else [1, 2][t % 2]
        if L == 12 and ed == 3 and t % 8 != 3:
            ed = 2
        keys = sorted({pack(rseq(rng, L)) for _ in range(int(rng.integers(1, 9)))})
        base = unpack(keys[int(rng.integers(len(keys)))], L)
        wstr = (mutate(rng, base, int(rng.integers(0, ed + 1))) + rseq(rng, 8))[:L]
        post = rseq(rng, 10 if bc else ed + 3)
        bail = None if t % 2 == 0 else int(rng.integers(1, 3))
        off = int(rng.integers(-2, 3))
        params = vm.new_object(params_cls, init=False)                   # field holder only: the testers read three nested fields
        if bc:
            bp = vm.new_object(vm.load("com/rw/parameters/BarcodeParameters"), init=False)
            allk = sorted(set(keys) | {pack((mutate(rng, wstr, int(rng.integers(0, 3))) + rseq(rng, 4))[:L]) for _ in range(4)})
            empk = sorted({pack((mutate(rng, wstr, int(rng.integers(1, 3))) + rseq(rng, 4))[:L]) for _ in range(3)})
            bp.f["cell_BC_bailout_after_ED"] = bail
            bp.f["maxEDtoCheckBCAll10xBCs"], bp.f["maxEDtoCheckBCEmptyDrops"] = 3, 2
            bp.f["checkAllassignedBarcodes"], bp.f["checkEmptyDrops"] = 1, 1
            params.f["barcodes"] = bp
            tester = vm.construct(BCT, "(Lcom/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams;Lcom/rw/nuc/reads/Illumina/All10xselectedCells;"
                                       "Lcom/rw/nuc/reads/Illumina/EmptyDropBarcodes;Lcom/rw/nuc/reads/Illumina/BarcodesMap;I"
                                       "Lcom/rw/nuc/encoding/onebyte/NucleicAcidInmutableOneBytePerBase;I)V",
                                  params, J.PySet(allk), J.PySet(empk), J.PySet(keys), ed, onebyte(vm, post), L)
        else:
            up = vm.new_object(vm.load("com/rw/parameters/UMIparameters"), init=False)
            up.f["umi_bailout_afterED"] = bail
            params.f["umis"] = up
            allk = empk = []
            tester = vm.construct(UMIT, "(ILcom/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams;"
                                        "Lcom/rw/nuc/reads/Illumina/byGene/IlluminaOneGeneOneCellData;Z"
                                        "Lcom/rw/nuc/encoding/onebyte/NucleicAcidInmutableOneBytePerBase;ZI)V",
                                  ed, params, J.PySet(keys), 1, onebyte(vm, post), 0, L)
        seq = vm.construct(T2, "(JI)V", J.L(pack(wstr)), L)
        lst = vm.call_virtual(tester, "matchesSeqEditDistance", "(Lcom/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase;I)Ljava/util/ArrayList;", seq, off)
        res = [(int(o.f["sequence"]) & M64, o.f["nSubstitutions"], o.f["nInsertions"], o.f["nDeletions"], o.f["startOffsetFromPredicted"],
                o.f["findingErrorFlag"]) for o in lst.v]
        cases.append(dict(bc=bc, L=L, ed=ed, keys=keys, allk=allk, empk=empk, w=pack(wstr), post=post, bail=-1 if bail is None else bail, off=off, res=res))
    return cases


def lev_cases(vm, rng, n):
    lev = vm.construct(LEV, "(Ljava/lang/Integer;)V", 4)
    rows = []
    for t in range(n):
        a = rng.choice([1, 2, 4, 8, 15], 12, p=[.24, .24, .24, .24, .04]).astype(np.uint8)
        b = a.copy()
        for _ in range(int(rng.integers(0, 7))):
            op, p = int(rng.integers(3)), int(rng.integers(12))
            if op == 0:
                b[p] = rng.choice([1, 2, 4, 8])
            elif op == 1:
                b = np.concatenate([b[:p], [rng.choice([1, 2, 4, 8])], b[p:]])[:12]
            else:
                b = np.concatenate([b[:p], b[p + 1:], [rng.choice([1, 2, 4, 8])]])
        d = vm.call_virtual(lev, "apply", "([B[B)Ljava/lang/Integer;", barr(a), barr(b))
        rows.append(np.concatenate([a, b, [np.uint8(d & 0xFF)]]))
    return np.array(rows, dtype=np.uint8)


def best9_cases(vm, rng, n, umi_len=12):
    """pair of reads -> 3 x 3 thresholded distances (LevenshteinDistance.apply on the reference's bytecode; the window slicing and the
    equal-bytes shortcut of lambda$static$7, ClusteringEditDistanceBase.java:L316-L343, are done here) -> new ClusteringEditDistanceBase(eds):
    best-of-9 in the reference's visiting order, BestEditDistance packing and its transposed copy"""
    CED = "com/rw/clustering/ClusteringEditDistanceBase"
    lev = vm.construct(LEV, "(Ljava/lang/Integer;)V", 4)
    rows = []
    for t in range(n):
        a = rng.choice([1, 2, 4, 8, 15], umi_len + 2, p=[.245, .245, .245, .245, .02]).astype(np.uint8)
        b = a.copy() if t % 4 else rng.choice([1, 2, 4, 8], umi_len + 2).astype(np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            op, p = int(rng.integers(3)), int(rng.integers(umi_len + 2))
            if op == 0:
                b[p] = rng.choice([1, 2, 4, 8])
            elif op == 1:
                b = np.concatenate([b[:p], [rng.choice([1, 2, 4, 8])], b[p:]])[:umi_len + 2]
            else:
                b = np.concatenate([b[:p], b[p + 1:], [rng.choice([1, 2, 4, 8])]])
        if t % 9 == 0:
            b = np.roll(a, 1 if t % 2 else -1)                      # a pure shift: best at a non-central position
        eds = J.JArr("L", 3, None)
        for i in range(3):
            row = J.JArr("B", 3, 0)
            for j in range(3):
                s1, s2 = a[i:i + umi_len], b[j:j + umi_len]
                if (s1 == s2).all():
                    row.a[j] = 0
                else:
                    d = vm.call_virtual(lev, "apply", "([B[B)Ljava/lang/Integer;", barr(s1), barr(s2))
                    row.a[j] = 5 if d == -1 else d
            eds.a[i] = row
        o = vm.construct(CED, "([[B)V", eds)
        best = o.f["bestEditDistance"]
        tr = vm.call_virtual(best, "getTransposedCopy", "()Lcom/rw/clustering/ClusteringEditDistanceBase$BestEditDistance;")
        rows.append(np.concatenate([a, b]).astype(np.int64).tolist() + [best.f["ed"], tr.f["ed"]])
    eq = vm.load(CED).statics["equalityEditDistance"].f["bestEditDistance"].f["ed"]
    return np.array(rows, dtype=np.int64), eq


def umi_pair_cases(vm, rng, n, umi_len=12):
    """calcEditDistances itself (ClusteringEditDistanceBase.lambda$static$7, java:L297-L350): two reads with their X= mini-sequences and the
    barcode end on the stranded mini-sequence -> window slicing (getSeqRevComp / getSeq, getSubSequence(bcEnd + 1 + i, umi_length)), the
    equal-bytes shortcut, the nine thresholded distances, best-of-9 and packing.  Only getCellBCendOnStrandedShortTestedSeq (the mapping
    of the read-name bcEnd onto the mini-sequence) is replaced by the value given here."""
    CED = "com/rw/clustering/ClusteringEditDistanceBase"
    ONR = "com/rw/umifinder/reads/nanopore/OneNanoporeResult"
    vm.init_class(vm.load(CED))
    onr = vm.load(ONR)
    vm.init_class(onr)
    comp = str.maketrans("ACGTN", "TGCAN")
    rows = []
    for t in range(n):
        five = t % 4 == 3
        params = bare(vm, "com/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams")
        up = bare(vm, "com/rw/parameters/UMIparameters")
        up.f["umi_length"] = umi_len
        params.f["umis"] = up
        st = vm.load("com/rw/parameters/ParametersMainBase$SCANTYPE")
        vm.init_class(st)
        params.f["scantype"] = st.statics["FIVEP_BARCODE" if five else "THREEP_BARCODE"]
        umi = rseq(rng, umi_len)
        recs, ends, strs = [], [], []
        for k in range(2):
            u = umi if k == 0 or t % 5 == 0 else mutate(rng, umi, int(rng.integers(0, 4)))
            if t % 7 == 0 and k == 1:
                u = u[:3] + "N" + u[4:]
            lead = int(rng.integers(3, 9))
            stranded = rseq(rng, lead) + u + rseq(rng, 24)                    # ... barcode end | UMI | polyA side
            bc_end = lead + (int(rng.integers(-1, 2)) if t % 3 == 0 else 0)    # 1-based end of the barcode on the stranded mini-sequence
            x = stranded if five else stranded[::-1].translate(comp)           # what the read name carries (read orientation)
            sd = vm.new_object(vm.load("com/rw/umifinder/reads/nanopore/NanoporeRead$ReadScanData"))
            vm.call_virtual(sd, "setSeq", "(Ljava/lang/CharSequence;)V", x)
            nr = bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead")
            nr.f["readScanData"] = J.JNative("com/google/common/base/Optional", (sd,))
            r = bare(vm, ONR)
            r.f["nanoporeRead"] = nr
            recs.append(r); ends.append(bc_end); strs.append(stranded)
        table = {id(recs[0]): ends[0], id(recs[1]): ends[1]}
        onr.statics["getCellBCendOnStrandedShortTestedSeq"] = J.JNative("pyfunc", lambda r, p: J.JNative("java/util/Optional", (table[id(r)],)))
        o = vm.invoke_exact(CED, "lambda$static$7", "(L%s;L%s;Lcom/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams;)L%s;" % (ONR, ONR, CED),
                            [recs[0], recs[1], params])
        rows.append((strs[0], strs[1], ends[0], ends[1], int(five), o.f["bestEditDistance"].f["ed"]))
    return rows


PKG = "com/rw/nanoporereadscanner/"


def bare(vm, name):
    return vm.new_object(vm.load(name), init=False)


def make_parser(vm, keys, ranks, ed, pm, three_prime, L=16):
    """a Parser whose fields hold just what assignBarcode reads: parameters (testPlusMinusPos, assignCellBCwithEditDistance, cell_bc_length,
    scantype), hashMapForBCfinding (the search set + CountsRank.rank) and the assignedBarcodes2ndPass map"""
    P = bare(vm, PKG + "analyzers/Parser")
    params, rsp, bp = bare(vm, PKG + "parameters/ParametersReadScannerApp"), bare(vm, "com/rw/parameters/ReadScannerParameters"), bare(vm, "com/rw/parameters/BarcodeParameters")
    rsp.f["testPlusMinusPos"] = pm
    rsp.f["assignCellBCwithEditDistance"] = J.JNative("com/google/common/base/Optional", (ed,))
    bp.f["cell_bc_length"] = L
    st = vm.load("com/rw/parameters/ParametersMainBase$SCANTYPE")
    vm.init_class(st)
    params.f["readScannerParameters"], params.f["barcodes"] = rsp, bp
    params.f["scantype"] = st.statics["THREEP_BARCODE" if three_prime else [k for k in st.statics if k.startswith("FIVEP")][0]]
    P.f["parameters"] = params
    bm = bare(vm, PKG + "WorkerReadscanner$BarcodesMapForBCfinding")
    bm.native = {}
    for k, r in zip(keys, ranks):
        cr = bare(vm, PKG + "WorkerReadscanner$CountsRank")
        cr.f["rank"] = int(r)
        bm.native[int(k)] = cr
    P.f["hashMapForBCfinding"] = bm
    P.f["assignedBarcodes2ndPass"] = J.JNative("java/util/concurrent/ConcurrentHashMap", {})
    return P


def run_assign(vm, P, read, adapterpos):
    """Parser.assignBarcode(fq) on the reference's bytecode.  Returns the BarcodeResult fields (None = nothing assigned) or 'EXC:<class>'"""
    fq = bare(vm, PKG + "readerwriter/FastqRecordExt")
    fq.f["strandedSequence"] = read
    sr = vm.new_object(vm.load(PKG + "readerwriter/ReadScanResult"))
    fq.f["scanResult"] = sr
    ar = vm.call_virtual(sr, "getAdapterresultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$Adapterresult;")
    ar.f["end"] = adapterpos
    try:
        vm.invoke_exact(PKG + "analyzers/Parser", "assignBarcode", "(L" + PKG + "readerwriter/FastqRecordExt;)V", [P, fq])
    except J.JavaThrow as ex:
        return "EXC:" + ex.cls
    br = [v for k, v in sr.f.items() if isinstance(v, J.JObj) and v.cls.name.endswith("BarcodeResult")]
    if not br or br[0].f.get("barcodeseq") is None:
        return None
    b = br[0].f
    rank = b.get("rank")
    rank = rank.v[0] if isinstance(rank, J.JNative) and rank.v else (rank if isinstance(rank, int) else -1)
    return (int(b["barcodeseq"].f["sequence"]) & M64, b["editDistance"], b["editDistanceSecondBest"], b["start"], b["end"], rank, int(sr.f["flag"]) & M64)


def assign_cases(vm, rng, n_cases):
    """whole reads through Parser.assignBarcode: 3' and 5' geometry, hits planted at several offsets / ED levels (ambiguity across offsets and
    levels), N in and around the window, windows running off the read"""
    L = 16
    comp = str.maketrans("ACGT", "TGCA")
    cases = []
    for t in range(n_cases):
        tp = t % 4 != 3
        ed = 1 if t % 5 == 0 else 2
        pm = 2 if t % 6 else 1
        read = rseq(rng, 60)
        ap = int(rng.integers(24, 36)) if t % 11 else int(rng.integers(14, 22))        # adapter end, 1-based; small values run off the read (3')
        if not tp and t % 11 == 0:
            ap = int(rng.integers(38, 46))
        keys = {pack(rseq(rng, L)) for _ in range(30)}
        for _ in range(int(rng.integers(1, 5))):                                        # plant barcodes near windows at random offsets
            o = int(rng.integers(-pm, pm + 1))
            if tp:
                b, e = ap - L - 1 + o, ap - 1 + o
                w = read[b:e][::-1].translate(comp) if 0 <= b and e <= len(read) else rseq(rng, L)
            else:
                b = ap + o
                w = read[b:b + L] if b + L <= len(read) else rseq(rng, L)
            if len(w) == L:
                keys.add(pack((mutate(rng, w, int(rng.integers(0, ed + 1))) + rseq(rng, 4))[:L]))
        if t % 9 == 0:
            p = int(rng.integers(max(0, ap - 24), min(60, ap + 22)))
            read = read[:p] + "N" + read[p + 1:]
        keys = sorted(keys)
        ranks = list(range(1, len(keys) + 1))
        P = make_parser(vm, keys, ranks, ed, pm, tp)
        res = run_assign(vm, P, read, ap)
        counts = {int(k): (v.f["counts"].v[0], {int(e): c.v[0] for e, c in v.f["edCounts"].v.items()}) for k, v in P.f["assignedBarcodes2ndPass"].v.items()}
        cases.append(dict(read=read, ap=ap, tp=tp, ed=ed, pm=pm, keys=keys, res=res, counts=counts))
    return cases


def find_umi_cases(vm, rng, n):
    """IlluminaUMIanalyzer.findUMI as a whole on the reference's bytecode (fixed edit distance path, java:L56-L58): the offset loop with its
    window / post-sequence geometry, one UMInucTwoBitPerBaseEDtester per offset, getBestAndSecondBCorUMI (sorted().distinct(), both
    NeedlemanWunsch alignments, MORE_THAN_ONE_MATCH) and the read positions.  Output: found, best and second entry, flag value, positions."""
    U = "com/rw/umifinder/"
    comp = str.maketrans("ACGTN", "TGCAN")
    st = vm.load("com/rw/parameters/ParametersMainBase$SCANTYPE")
    vm.init_class(st)
    cases = []
    for t in range(n):
        ed, pm = [1, 2, 2, 2][t % 4], [1, 2][t % 2]
        bail = None if t % 3 else 1
        umi = rseq(rng, 12, "AAAGCT" if t % 5 == 0 else "AGCT")
        umis = {pack(umi)} | {pack(rseq(rng, 12)) for _ in range(int(rng.integers(0, 6)))}
        for _ in range(int(rng.integers(0, 3))):                                 # close relatives: a second-best match / ambiguity
            umis.add(pack((mutate(rng, umi, int(rng.integers(1, 3))) + rseq(rng, 3))[:12]))
        lead = int(rng.integers(4, 9))
        obs = umi if t % 4 == 0 else mutate(rng, umi, int(rng.integers(0, ed + 2)))
        stranded = rseq(rng, lead) + obs + rseq(rng, 26)
        bc_end = lead + int(rng.integers(-pm, pm + 1)) * (t % 2)                 # predicted barcode end off by up to +-pm
        if t % 13 == 12:
            stranded = stranded[:bc_end + 12 + 2]                                # read ends inside the post sequence: padded with 'A' (java:L118-L124)
        params = bare(vm, U + "parameters/ParametersBarcodeUMiFinderAppParams")
        up = bare(vm, "com/rw/parameters/UMIparameters")
        up.f.update(maxUMIfalseAssignmentPcnt=None, umi_editdistance=ed, incrementUMI_ED_ifFewUmis=0, umi_posplusminus=pm, umi_length=12,
                    simulateRandomUmis=0, umi_bailout_afterED=bail, maxUMIsforincrementUMI_ED=50)
        params.f["umis"] = up
        params.f["scantype"] = st.statics["THREEP_BARCODE"]
        nd = bare(vm, "com/rw/nanopore/analyzers/parameters/NeedlemanParameters")
        nd.f["umi"] = vm.construct("com/rw/nuc/alignment/needleman/NeedlemanScores", "()V")
        params.f["needleman"] = nd
        ana = bare(vm, U + "analyzers/IlluminaUMIanalyzer")
        ana.f["parameters"], ana.f["scanStats"] = params, bare(vm, U + "scanstats/ScanStats")
        sd = vm.new_object(vm.load(U + "reads/nanopore/NanoporeRead$ReadScanData"))
        vm.call_virtual(sd, "setSeq", "(Ljava/lang/CharSequence;)V", stranded[::-1].translate(comp))
        nr = bare(vm, U + "reads/nanopore/NanoporeRead")
        nr.f["readScanData"] = J.JNative("com/google/common/base/Optional", (sd,))
        r = bare(vm, U + "reads/nanopore/OneNanoporeResult")
        r.f["nanoporeRead"] = nr
        m = bare(vm, "com/rw/nanopore/analyzers/Match")
        m.f["endposInTestedSeq"], m.f["endPosRead"] = J.JNative("java/util/Optional", (bc_end,)), J.JNative("java/util/Optional", (500,))
        r.f["cellBCMatch"], r.f["umiMatch"], r.f["umiFindingFlagValue"] = J.JNative("java/util/Optional", (m,)), J.JNative("java/util/Optional", ()), 0
        try:
            found = vm.call_virtual(ana, "findUMI", "(L%sreads/nanopore/OneNanoporeResult;Lcom/rw/nuc/reads/Illumina/byGene/IlluminaOneGeneOneCellData;)Z" % U,
                                    r, J.PySet(umis))
        except J.JavaThrow as ex:
            cases.append(dict(stranded=stranded, bc_end=bc_end, umis=sorted(umis), ed=ed, pm=pm, bail=-1 if bail is None else bail, exc=ex.cls, row=[0] * 16))
            continue
        ent = lambda o: [int(o.f["sequence"]) & M64, o.f["nSubstitutions"], o.f["nInsertions"], o.f["nDeletions"], o.f["startOffsetFromPredicted"]]
        row = [int(found), r.f["umiFindingFlagValue"]]
        um = r.f["umiMatch"].v[0] if r.f["umiMatch"].v else None
        best = r.f.get("umi")
        row += ent(best) if isinstance(best, J.JObj) else [0, 0, 0, 0, 0]
        sec = um.f["secondBestMatch"] if um is not None else None
        secnode = sec.f["alignment"].f.get("mutSeq") if sec is not None and isinstance(sec.f.get("alignment"), J.JObj) else None
        row += [int(sec is not None)]
        row += [um.f["startposRead"].v[0], um.f["endPosRead"].v[0]] if um is not None and um.f["startposRead"].v else [0, 0]
        row += [um.f["nMismatchDiffBestvsSecondBest"] if um is not None and um.f["nMismatchDiffBestvsSecondBest"] is not None else -99]
        row += [len(sec.f["alignment"].f["read"]) if secnode is None and sec is not None else 0]
        sec_read = sec.f["alignment"].f["match"].replace("-", "") if sec is not None else ""
        cases.append(dict(stranded=stranded, bc_end=bc_end, umis=sorted(umis), ed=ed, pm=pm, bail=-1 if bail is None else bail, exc="", row=row, second=sec_read))
    return cases


def test_barcodes_cases(vm, rng, n):
    """IlluminaBarcodeAnalyzer.testBarcodes (one gene) + getBestAndSecondBCorUMI(CELLBC) on the reference's bytecode: the BC-flavour offset loop
    (window = 16 bases after the nbasesOfAdapterSeqInReadname adapter bases of the X= mini-sequence, 10 post bases), BCnucTwoBitPerBaseEDtester with
    the gene / all-passed / empty-drop lists and the bailout, then the sort with scoreWhereFound + distinct and the Needleman alignments"""
    U = "com/rw/umifinder/"
    IBA = U + "analyzers/IlluminaBarcodeAnalyzer"
    comp = str.maketrans("ACGTN", "TGCAN")
    st = vm.load("com/rw/parameters/ParametersMainBase$SCANTYPE")
    vm.init_class(st)
    sw = vm.load("com/rw/nanopore/analyzers/AnalyzerBase$ScanningWhat")
    vm.init_class(sw)
    cases = []
    for t in range(n):
        ed, pm = [1, 2][t % 2], [1, 2, 2][t % 3]
        bail = [None, 1, 2][t % 3]
        bc = rseq(rng, 16, "AAAGCT" if t % 5 == 0 else "AGCT")
        gene = {pack(rseq(rng, 16)) for _ in range(int(rng.integers(0, 12)))}
        if t % 4:
            gene.add(pack(bc))
        allk = set(gene) | {pack(rseq(rng, 16)) for _ in range(10)} | {pack((mutate(rng, bc, int(rng.integers(0, 3))) + rseq(rng, 3))[:16]) for _ in range(3)}
        empk = {pack(rseq(rng, 16)) for _ in range(10)} | {pack((mutate(rng, bc, int(rng.integers(1, 3))) + rseq(rng, 3))[:16]) for _ in range(2)}
        obs = mutate(rng, bc, int(rng.integers(0, ed + 2)))
        lead = 3 + (int(rng.integers(-pm, pm + 1)) if t % 2 else 0)
        stranded = rseq(rng, max(lead, 0)) + obs + rseq(rng, 30)
        params = bare(vm, U + "parameters/ParametersBarcodeUMiFinderAppParams")
        bp, rsp = bare(vm, "com/rw/parameters/BarcodeParameters"), bare(vm, "com/rw/parameters/ReadScannerParameters")
        bp.f.update(bc_posplusminus=pm, simulateRandomBCs=0, cell_bc_length=16, cell_BC_bailout_after_ED=bail, maxEDtoCheckBCAll10xBCs=3,
                    maxEDtoCheckBCEmptyDrops=2, checkAllassignedBarcodes=1, checkEmptyDrops=1)
        rsp.f["nbasesOfAdapterSeqInReadname"] = 3
        params.f["barcodes"], params.f["readScannerParameters"], params.f["scantype"] = bp, rsp, st.statics["THREEP_BARCODE"]
        ill = J.JNative("ParsedIlluminaData")
        ill.f["all10xselectedCells"] = J.PySet(allk)
        params.f["illuminaData"] = J.JNative("com/google/common/base/Optional", (ill,))
        scores = vm.construct("com/rw/nuc/alignment/needleman/NeedlemanScores", "()V")
        sd = vm.new_object(vm.load(U + "reads/nanopore/NanoporeRead$ReadScanData"))
        vm.call_virtual(sd, "setSeq", "(Ljava/lang/CharSequence;)V", stranded[::-1].translate(comp))
        nr = bare(vm, U + "reads/nanopore/NanoporeRead")
        nr.f["readScanData"] = J.JNative("com/google/common/base/Optional", (sd,))
        r = bare(vm, U + "reads/nanopore/OneNanoporeResult")
        r.f["nanoporeRead"], r.f["bcFindingFlagValue"] = nr, 0
        ana = bare(vm, IBA)
        ana.f["parameters"], ana.f["scanStats"], ana.f["oneNanoporeResult"] = params, bare(vm, U + "scanstats/ScanStats"), r
        ogd = vm.construct(IBA + "$OneGeneOrRegionData", "(Lcom/rw/nuc/reads/Illumina/BarcodesMap;Ljava/lang/String;Z)V", J.PySet(gene, J.PySet(empk)), "GENE1", 0)
        ogd.f["maxEDdyn"] = ed
        row = dict(stranded=stranded, gene=sorted(gene), allk=sorted(allk), empk=sorted(empk), ed=ed, pm=pm, bail=-1 if bail is None else bail, exc="")
        try:
            e = vm.invoke_exact(IBA, "testBarcodes", "(Ljava/util/List;)Ljava/util/Map$Entry;", [ana, J.JNative("java/util/ArrayList", [ogd])])
            ent = lambda o: [int(o.f["sequence"]) & M64, o.f["nSubstitutions"], o.f["nInsertions"], o.f["nDeletions"], o.f["startOffsetFromPredicted"], o.f["findingErrorFlag"]]
            if e is None:
                row.update(n_raw=0, best=[0] * 6, second=[0] * 6, n_distinct=0, min_err_gene=2147483647)
            else:
                lst = e.v[1].v
                ge = [o.f["nSubstitutions"] + o.f["nInsertions"] + o.f["nDeletions"] for o in lst if o.f["findingErrorFlag"] & 512]
                m = bare(vm, "com/rw/nanopore/analyzers/Match")
                best = vm.invoke_exact(U + "analyzers/IlluminaBarcodeUMIAnalyzerBase", "getBestAndSecondBCorUMI",
                                       "(L%sreads/nanopore/OneNanoporeResult;Ljava/util/List;Ljava/util/Optional;Lcom/rw/nanopore/analyzers/AnalyzerBase$ScanningWhat;"
                                       "Lcom/rw/nuc/alignment/needleman/NeedlemanScores;)Lcom/rw/nuc/encoding/TwoBit/NucTwoBitPerBaseWithErrors;" % U,
                                       [r, e.v[1], J.JNative("java/util/Optional", (m,)), sw.statics["CELLBC"], scores])
                sec = m.f["secondBestMatch"]
                sec_seq = pack(sec.f["alignment"].f["match"].replace("-", "")) if sec is not None else 0
                row.update(n_raw=len(lst), best=ent(best), second=[sec_seq, 0, 0, 0, 0, 0], n_distinct=2 if sec is not None else 1, min_err_gene=min(ge) if ge else 2147483647,
                           mismatch_diff=-99 if m.f["nMismatchDiffBestvsSecondBest"] is None else int(m.f["nMismatchDiffBestvsSecondBest"]), bc_flag=int(r.f["bcFindingFlagValue"]))
        except J.JavaThrow as ex:
            row.update(exc=ex.cls, n_raw=0, best=[0] * 6, second=[0] * 6, n_distinct=0, min_err_gene=2147483647)
        cases.append(row)
    return cases


TAGS = dict(paStartPrefix="PS=", paEndPrefix="PE=", adapterPosPrefix="AE=", tsoPosPrefix="T=", seqPrefix="X=", qvPrefix="Q=", barcodeSeqPrefix="bc=",
            barcodeEdPrefix="ed=", barcodeEdSecondaryPrefix="ed_sec=", barcodeStartPrefix="bcStart=", barcodeEndPrefix="bcEnd=", barcodeRankPrefix="rk=")


def read_name_cases(vm, rng, n):
    """FastqRecordExt.getScanDatFromReadName (the assignumis side of the read-name format) on the reference's bytecode: the two README
    examples and generated names (missing tags, T=, sp2 prefix, ed above the assignumis limit).  Output: the parsed fields."""
    F = PKG + "readerwriter/FastqRecordExt"
    names = ["b5c7-read_FWD_PS=566_PE=590_AE=619_bc=TCCGATCGTGCCAAGA_ed=0_ed_sec=2147483647_bcStart=618_bcEnd=603_rk=2987_X=AAAAAAAAAAAATGGCGTGTATTGTCTTGGCACGATCGGAAGA_Q=27.1",
             "sp2_REV_PS=1257_PE=1305_AE=1327_T=40_bc=GAGTGAGGTTGGGTAG_ed=1_ed_sec=2147483647_bcStart=1326_bcEnd=1311_rk=3883_X=AAAAAAAAAAACAAACCAAGTAACCAACCCAACCTCACTCAGA_Q=15.9",
             "no_tags_at_all", "r_FWD_PS=5_PE=9_"]
    for t in range(n):
        nm = "read%d" % t + ("_REV" if t % 2 else "_FWD") + "_"
        if t % 3:
            nm += "PS=%d_PE=%d_" % (rng.integers(10, 900), rng.integers(10, 900))
        nm += "AE=%d_" % rng.integers(20, 2000)
        if t % 4 == 0:
            nm += "T=%d_" % rng.integers(1, 99)
        if t % 5:
            nm += "bc=%s_ed=%d_" % (rseq(rng, 16), rng.integers(0, 4))
            if t % 7:
                nm += "ed_sec=%d_" % rng.choice([1, 2, 3, 2147483647])
            nm += "bcStart=%d_bcEnd=%d_" % (rng.integers(20, 2000), rng.integers(20, 2000))
            if t % 6:
                nm += "rk=%d_" % rng.integers(1, 9000)
        nm += "X=%s_Q=%.1f" % (rseq(rng, 43), rng.uniform(5, 40))
        if t % 8 == 7:
            pass                                                   # no trailing '_': the base-36 read-id parse throws (like the README's trimmed examples)
        elif t % 2 == 0:
            nm += "_"                                              # what scanfastq writes: '_' after Q= (FastqRecordExt.java:L271) ...
        else:
            nm += "_" + vm.call_static(F + "$NumberToAndFromAscii", "convertInt", "(I)Ljava/lang/String;", int(rng.integers(0, 10 ** 8)))   # ... + the base-36 read id
        names.append(nm)
    out = []
    for lim in (None, 1):
        params = bare(vm, "com/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams")
        rsp, bp = bare(vm, "com/rw/parameters/ReadScannerParameters"), bare(vm, "com/rw/parameters/BarcodeParameters")
        for k, v in TAGS.items():
            rsp.f[k] = v
        bp.f["cellBC_editdistance"] = lim
        params.f["readScannerParameters"], params.f["barcodes"] = rsp, bp
        for nm in names:
            sup = J.JNative("pyfunc", lambda: vm.new_object(vm.load(PKG + "readerwriter/ReadScanResult")))
            try:
                r = vm.call_static(F, "getScanDatFromReadName", "(Ljava/lang/String;Lcom/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams;"
                                                                "Ljava/util/function/Supplier;)Lcom/google/common/base/Optional;", nm, params, sup)
            except J.JavaThrow as ex:
                out.append((nm, -1 if lim is None else lim, "EXC:" + ex.cls.split("$")[-1].split("/")[-1]))
                continue
            if not r.v:
                out.append((nm, -1 if lim is None else lim, "ABSENT"))
                continue
            sr = r.v[0]
            d = {}
            sub = lambda o: o.f if isinstance(o, J.JObj) else {}
            for k, v in sr.f.items():
                if isinstance(v, J.JObj):
                    cn = v.cls.name.split("$")[-1]
                    for kk, vv in v.f.items():
                        if isinstance(vv, (int, float)) and not isinstance(vv, bool):
                            d[cn + "." + kk] = vv
                        elif isinstance(vv, J.JObj) and "sequence" in vv.f:
                            d[cn + "." + kk] = int(vv.f["sequence"]) & M64
                        elif isinstance(vv, J.JNative) and vv.name.endswith("Optional") and vv.v:
                            d[cn + "." + kk] = vv.v[0]
            d["read_id"] = sr.f.get("read_id")
            for k, v in sr.f.items():
                if isinstance(v, float):
                    d[k] = float(np.float32(v))
            if isinstance(sr.f.get("seq"), J.JObj):
                d["seq_len"] = len(sr.f["seq"].f["naData"].a)
            out.append((nm, -1 if lim is None else lim, repr(sorted(d.items()))))
    return out


def write_name_cases(vm, rng, n):
    """FastqRecordExt.getRecordForWriting (the scanfastq side of the read-name format) on the reference's bytecode: the read name it builds
    for a passed read from polyA / adapter / TSO / barcode results, the X= slice of the stranded read and the mean quality"""
    F = PKG + "readerwriter/FastqRecordExt"
    fl = vm.load(PKG + "stats/ReadFlags$Flags")
    vm.init_class(fl)
    flagv = lambda k: int(vm.call_virtual(fl.statics[k], "getValue", "()J"))
    rsp = bare(vm, "com/rw/parameters/ReadScannerParameters")
    for k, v in TAGS.items():
        rsp.f[k] = v
    rsp.f["nbasesOfAdapterSeqInReadname"] = 3
    rsp.f["trimFastq"] = 0
    out = []
    for t in range(n):
        rev, five = bool(t % 2), t % 5 == 4
        stranded = rseq(rng, 160)
        quals = "".join(chr(33 + int(q)) for q in rng.integers(3, 42, 160))
        ae = int(rng.integers(45, 110)) if t % 9 else int(rng.integers(20, 44))          # small adapter ends: "Beginrange inconsistent" -> no X= / Q=
        fq = bare(vm, F)
        fq.f["strandedSequence"] = stranded
        sr = vm.new_object(vm.load(PKG + "readerwriter/ReadScanResult"))
        sr.f["flag"] = J.L(flagv("PASSED_REV" if rev else "PASSED_FWD"))
        fq.f["scanResult"] = sr
        kw = dict(adapter_end=ae)
        vm.call_virtual(sr, "getAdapterresultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$Adapterresult;").f["end"] = ae
        if t % 3:
            pa = vm.call_virtual(sr, "getPolyAResultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$PolyAResult;")
            kw["polya_start"], kw["polya_end"] = int(rng.integers(1, 40)), int(rng.integers(40, 44))
            pa.f["start"], pa.f["end"] = kw["polya_start"], kw["polya_end"]
        if t % 4 == 0:
            kw["tso_end"] = int(rng.integers(1, 90))
            vm.call_virtual(sr, "getTSOresultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$TSOresult;").f["end"] = kw["tso_end"]
        if t % 6:
            br = vm.call_virtual(sr, "getBarcodeResultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$BarcodeResult;")
            kw.update(bc=rseq(rng, 16), ed=int(rng.integers(0, 3)), bc_start=int(rng.integers(1, 150)), bc_end=int(rng.integers(1, 150)))
            vm.call_virtual(br, "setBarcodeseq", "(Ljava/lang/String;)V", kw["bc"])
            br.f["editDistance"], br.f["start"], br.f["end"] = kw["ed"], kw["bc_start"], kw["bc_end"]
            if t % 7:
                kw["ed_second"] = int(rng.choice([1, 2, 2147483647]))
                br.f["editDistanceSecondBest"] = kw["ed_second"]
            if t % 5:
                kw["rank"] = int(rng.integers(1, 5000))
                vm.call_virtual(br, "setRank", "(I)V", kw["rank"])
        rid = None if t % 3 == 0 else int(rng.integers(0, 10 ** 7))
        # the superclass htsjdk.samtools.fastq.FastqRecord is outside the jars: its three getters are answered from here
        vm.fastq_fields = {"getReadName": "read%d some description" % t, "getBaseQualityString": quals[::-1] if rev else quals, "getReadString": stranded,
                           "getBaseQualityHeader": "", "toString": "FastqRecord"}
        fq.f["reverseComplementQualities"] = J.JNative("pyfunc", lambda rec: quals)       # FastqRecordExt's own lambda reverses the qualities of a reversed read
        try:
            rec = vm.call_virtual(fq, "getRecordForWriting", "(Lcom/rw/parameters/ReadScannerParameters;ZLjava/lang/Integer;)Lhtsjdk/samtools/fastq/FastqRecord;",
                                  rsp, int(five), rid)
            name = rec.v[0]
        except J.JavaThrow as ex:                              # e.g. the X= slice starting at the first base: getMeanQV's skip(begin - 1) gets -1
            name = "EXC:" + ex.cls
        out.append(dict(name=name, stranded=stranded, quals=quals, rev=rev, five=five, rid=-1 if rid is None else rid, kw=repr(sorted(kw.items()))))
    return out


def exact_lookup_cases(vm, rng, n):
    """pass 1, UsedCellBCListGenerator$Worker.lambda$call$1 (UsedCellBCListGenerator.java:L206-L232) on the reference's bytecode: the window at
    the predicted position (3' reverse-complemented), the whitelist test, the per-barcode read counts"""
    G = PKG + "analyzers/UsedCellBCListGenerator"
    st = vm.load("com/rw/parameters/ParametersMainBase$SCANTYPE")
    vm.init_class(st)
    comp = str.maketrans("ACGT", "TGCA")
    out = []
    for tp in (True, False):
        params, bp = bare(vm, PKG + "parameters/ParametersReadScannerApp"), bare(vm, "com/rw/parameters/BarcodeParameters")
        bp.f["cell_bc_length"] = 16
        params.f["barcodes"], params.f["scantype"] = bp, st.statics["THREEP_BARCODE" if tp else [k for k in st.statics if k.startswith("FIVEP")][0]]
        gen, dbg, used = bare(vm, G), bare(vm, G + "$DebugInfo"), bare(vm, G + "$UsedBarcodesListData")
        for k in list(dbg.f):
            dbg.f[k] = J.JNative("java/util/concurrent/atomic/AtomicInteger", [0])
        used.f["unfilteredUsedBarcodeMap"] = J.JNative("it/unimi/dsi/fastutil/longs/Long2ObjectMap", {})
        used.f["recordCount"] = J.JNative("java/util/concurrent/atomic/AtomicInteger", [0])
        gen.f["params"], gen.f["debugInfo"], gen.f["barcodesUsedData"] = params, dbg, used
        wk = bare(vm, G + "$Worker")
        wk.f["this$0"] = gen
        whitelist = {pack(rseq(rng, 16)) for _ in range(60)}
        pred = J.JNative("pyfunc", lambda k: int(int(k) in whitelist))
        reads = []
        for t in range(n):
            read = rseq(rng, 70)
            ap = int(rng.integers(20, 50)) if t % 10 else int(rng.integers(5, 17))
            if t % 3:                                                                    # plant a whitelist barcode at the predicted position
                b = unpack(sorted(whitelist)[int(rng.integers(len(whitelist)))], 16)
                if tp and ap - 17 >= 0:
                    read = read[:ap - 17] + b[::-1].translate(comp) + read[ap - 1:]
                elif not tp:
                    read = (read[:ap] + b + read[ap + 16:])[:70]
            if t % 11 == 0:
                read = read[:ap - 5] + "N" + read[ap - 4:]
            fq = bare(vm, PKG + "readerwriter/FastqRecordExt")
            fq.f["strandedSequence"] = read
            sr = vm.new_object(vm.load(PKG + "readerwriter/ReadScanResult"))
            fq.f["scanResult"] = sr
            vm.call_virtual(sr, "getAdapterresultCreateIfNull", "()L" + PKG + "readerwriter/ReadScanResult$Adapterresult;").f["end"] = ap
            try:
                r = vm.invoke_exact(G + "$Worker", "lambda$call$1", "(L" + PKG + "readerwriter/FastqRecordExt;Ljava/util/function/Predicate;)Ljava/lang/Boolean;",
                                    [wk, fq, pred])
            except J.JavaThrow as ex:
                r = -1
            reads.append((read, ap, int(r)))
        counts = {k: v.v[0] for k, v in used.f["unfilteredUsedBarcodeMap"].v.items()}
        out.append(dict(tp=tp, whitelist=sorted(whitelist), reads=reads, counts=counts))
    return out


def getmaxed_cases(vm):
    """DynamicEditDistances.getmaxED on the reference's bytecode, fed with the reference's own tables (Jar/bcMaxEditDistances.xml,
    Jar/umiMaxEditDistances.xml, parsed here instead of through JAXB)"""
    import xml.etree.ElementTree as ET
    DED = "com/rw/parameters/DynamicEditDistances"
    rows = []
    for fn in ("bcMaxEditDistances.xml", "umiMaxEditDistances.xml"):
        tables = {}
        for lc in ET.parse(REF + "/" + fn).getroot().iter("dataForUMIlength"):
            L = int(lc.find("umiBCLength").text)
            for ec in lc.iter("dataForErr"):
                tables[(L, int(ec.find("errorpercent").text))] = {int(d.find("editDistance").text): int(d.find("maxBarcodes").text) for d in ec.iter("dataForED")}
        d = bare(vm, DED)
        entries = {}
        for (L, err), col in tables.items():
            le = entries.get(L)
            if le is None:
                le = entries[L] = bare(vm, DED + "$OneUMIBClengthEntry")
                le.native = {}
            ee = bare(vm, DED + "$OneErrPctEntry")
            ee.native = {k: J.L(v) for k, v in col.items()}
            le.native[err] = ee
        d.f["entries"] = J.JNative("java/util/HashMap", entries)
        for (L, err), col in sorted(tables.items()):
            for count in (1, 2, 3, 7, 20, 21, 100, 570, 2849, 2850, 26362, 100000, 600001):
                for pm in (0, 1, 2):
                    for cap in (None, 0, 2, 3):
                        try:
                            r = vm.call_virtual(d, "getmaxED", "(IIIILjava/lang/Integer;)I", count, pm, err, L, cap)
                        except J.JavaThrow:
                            r = -1
                        rows.append([L, err, count, pm, -1 if cap is None else cap, r] + [col.get(e, -1) for e in range(6)])
    return np.array(rows, dtype=np.int64)


def save_test_barcodes(tb):
    gk, go = flat(tb, "gene")
    ak, ao = flat(tb, "allk")
    ek, eo = flat(tb, "empk")
    np.savez_compressed(os.path.join(OUT, "ref_test_barcodes.npz"), stranded=np.array([c["stranded"] for c in tb]), gene=gk, gene_offsets=go, all_keys=ak, all_offsets=ao,
                        empty_keys=ek, empty_offsets=eo, ed=np.array([c["ed"] for c in tb], dtype=np.int32), pm=np.array([c["pm"] for c in tb], dtype=np.int32),
                        bail=np.array([c["bail"] for c in tb], dtype=np.int32), exc=np.array([c["exc"] for c in tb]), n_raw=np.array([c["n_raw"] for c in tb], dtype=np.int64),
                        best=np.array([c["best"] for c in tb], dtype=np.int64), second=np.array([c["second"] for c in tb], dtype=np.int64),
                        n_distinct=np.array([c["n_distinct"] for c in tb], dtype=np.int32), min_err_gene=np.array([c["min_err_gene"] for c in tb], dtype=np.int64),
                        mismatch_diff=np.array([c.get("mismatch_diff", -99) for c in tb], dtype=np.int32),      # nMismatchDiffBestvsSecondBest, -99 = no second
                        bc_flag=np.array([c.get("bc_flag", 0) for c in tb], dtype=np.int64))                     # bcFindingFlagValue after getBestAndSecondBCorUMI


CLUSTER_SIG = "clusterLocal(Ljava/util/Collection;Lcom/rw/clustering/DistanceMatrix;)Ljava/util/Optional;"


def cluster_cases(vm, rng, n_cases):
    """ClusterOne_MyClustering.clusterLocal (ClusterOne_MyClustering.java:L175-L219) on packed matrices: the reference's own streams,
    lambdas, DistanceMatrix.distanceNonReducedSet and BestEditDistance.getED run as bytecode; java.util.stream, HashSet, Collectors
    and fastutil's Int2ObjectOpenHashMap are shims (sequential pipelines; the map's iteration order is oracle/pyref.fastutil_key_order
    and is stored with every case, because the reference's own order for > 30 reads depends on how its parallel stream was split)."""
    out = []
    for i in range(n_cases):
        n = int(rng.integers(2, 40))
        p_close = float(rng.choice([0.05, 0.2, 0.5, 0.9]))
        ed_m = np.where(rng.random((n, n)) < p_close, rng.integers(0, 3, (n, n)), rng.integers(3, 6, (n, n))).astype(np.int64)
        if i % 4:
            ed_m = np.triu(ed_m, 1)
            ed_m = ed_m + ed_m.T
        if i % 5 == 0:                                       # blocks of identical reads: many equal neighbour counts
            lab = rng.integers(0, max(1, n // 4), n)
            ed_m = np.where(lab[:, None] == lab[None, :], 0, 5).astype(np.int64)
        np.fill_diagonal(ed_m, 0)
        packed = (ed_m | (rng.integers(0, 64, (n, n)) << 24)).astype(np.int32)
        ed = int(rng.choice([0, 1, 2, 2, 3]))
        member = np.ones(n, dtype=np.uint8) if i % 3 else (rng.random(n) < 0.7).astype(np.uint8)
        me = bare(vm, "com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering")
        me.f["ed"] = ed
        dm = bare(vm, "com/rw/clustering/DistanceMatrix")
        rows = J.JArr("[", n, None)
        for a in range(n):
            row = J.JArr("L", n, None)
            for v in range(n):
                cell = bare(vm, "com/rw/clustering/ClusteringEditDistanceBase")
                best = bare(vm, "com/rw/clustering/ClusteringEditDistanceBase$BestEditDistance")
                best.f["ed"] = int(packed[a, v])
                cell.f["bestEditDistance"] = best
                row.a[v] = cell
            rows.a[a] = row
        dm.f["distanceMatrix"] = rows
        indices = J.JNative("java/util/ArrayList", [int(x) for x in np.flatnonzero(member)])
        captured = {}
        orig_collect = vm.collect

        def spy(items, col, _orig=orig_collect):
            r = _orig(items, col)
            if isinstance(r.v, J.FastutilIntMap):
                captured["order"] = list(r.v.order())
            return r
        vm.collect = spy
        try:
            opt = vm.run(me.cls, CLUSTER_SIG, [me, indices, dm])
        finally:
            vm.collect = orig_collect
        label = np.full(n, -1, dtype=np.int32)               # cluster of read c = smallest member of its cluster, -1 = in no cluster
        if opt.v:
            for cl in opt.v[0].v.items():
                mem = sorted(int(x) for x in cl.v.items())
                for x in mem:
                    assert label[x] == -1
                    label[x] = mem[0]
        rank = np.full(n, 2 ** 30, dtype=np.int32)
        for r, k in enumerate(captured.get("order", [])):
            rank[k] = r
        out.append(dict(n=n, ed=ed, packed=packed, member=member, rank=rank, label=label, present=int(bool(opt.v))))
    return out


def save_cluster_cases(cc):
    off = np.cumsum([0] + [c["n"] for c in cc]).astype(np.int64)
    moff = np.cumsum([0] + [c["n"] ** 2 for c in cc]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "ref_cluster_local.npz"), job_offsets=off, out_offsets=moff,
                        packed=np.concatenate([c["packed"].ravel() for c in cc]), ed=np.array([c["ed"] for c in cc], dtype=np.int32),
                        member=np.concatenate([c["member"] for c in cc]), rank=np.concatenate([c["rank"] for c in cc]),
                        label=np.concatenate([c["label"] for c in cc]), present=np.array([c["present"] for c in cc], dtype=np.int32))


def flat(cases, key):
    off = np.cumsum([0] + [len(c[key]) for c in cases]).astype(np.int64)
    return np.array([k for c in cases for k in c[key]], dtype=np.uint64), off


def main():
    vm = J.VM(JARS + [REF + "/lib/commons-lang3-3.17.0.jar"])      # ImmutablePair (ClusterOne_MyClustering's lambdas) runs as bytecode too
    t0 = time.time()
    cc = cluster_cases(vm, np.random.default_rng(1717), 60)
    save_cluster_cases(cc)
    print("clusterLocal", len(cc), "jobs, with clusters", sum(c["present"] for c in cc), "reads in clusters", sum(int((c["label"] >= 0).sum()) for c in cc),
          "%.1fs" % (time.time() - t0), vm.n_insn, "bytecodes")
    if len(sys.argv) > 1 and sys.argv[1] == "cluster":
        return
    rng = np.random.default_rng(20261017)
    t0 = time.time()
    prim, seqs = primitives(vm, rng)
    np.savez_compressed(os.path.join(OUT, "ref_primitives.npz"), rows=prim, seq=np.array([x[0].ljust(16, "-") for x in seqs]),
                        seq_len=np.array([x[1] for x in seqs], dtype=np.int32), seq_hash=np.array([x[2] for x in seqs], dtype=np.uint64),
                        seq_revcomp=np.array([x[3] for x in seqs], dtype=np.uint64))
    print("primitives", prim.shape, "%.1fs" % (time.time() - t0), vm.n_insn, "bytecodes")

    lv = lev_cases(vm, rng, 400)
    np.savez_compressed(os.path.join(OUT, "ref_levenshtein.npz"), rows=lv)
    print("levenshtein", lv.shape, "d histogram", np.bincount(lv[:, 24].astype(np.int8).astype(int) + 1))

    b9, eq = best9_cases(vm, np.random.default_rng(99), 300)
    np.savez_compressed(os.path.join(OUT, "ref_best9.npz"), rows=b9, equality=np.int64(eq))
    print("best-of-9", b9.shape, "ED histogram", np.bincount(b9[:, 28] & 0xFFFFFF), "equality %#x" % eq)

    up = umi_pair_cases(vm, np.random.default_rng(4242), 200)
    np.savez_compressed(os.path.join(OUT, "ref_umi_pairs.npz"), s1=np.array([r[0] for r in up]), s2=np.array([r[1] for r in up]),
                        end1=np.array([r[2] for r in up], dtype=np.int32), end2=np.array([r[3] for r in up], dtype=np.int32),
                        five_prime=np.array([r[4] for r in up], dtype=np.int32), packed=np.array([r[5] for r in up], dtype=np.int64))
    print("calcEditDistances", len(up), "pairs, ED histogram", np.bincount(np.array([r[5] for r in up]) & 0xFFFFFF))

    rn = read_name_cases(vm, np.random.default_rng(31), 40)
    np.savez_compressed(os.path.join(OUT, "ref_read_names.npz"), name=np.array([r[0] for r in rn]), limit=np.array([r[1] for r in rn], dtype=np.int32),
                        parsed=np.array([r[2] for r in rn]))
    print("getScanDatFromReadName", len(rn), "names")

    fu = find_umi_cases(vm, np.random.default_rng(606), 48)
    keys, koff = flat(fu, "umis")
    np.savez_compressed(os.path.join(OUT, "ref_find_umi.npz"), stranded=np.array([c["stranded"] for c in fu]), bc_end=np.array([c["bc_end"] for c in fu], dtype=np.int32),
                        umis=keys, umi_offsets=koff, ed=np.array([c["ed"] for c in fu], dtype=np.int32), pm=np.array([c["pm"] for c in fu], dtype=np.int32),
                        bail=np.array([c["bail"] for c in fu], dtype=np.int32), exc=np.array([c["exc"] for c in fu]),
                        row=np.array([c["row"][:12] for c in fu], dtype=np.int64), second=np.array([c.get("second", "") for c in fu]))
    print("findUMI", len(fu), "reads, found", sum(c["row"][0] for c in fu), "with second", sum(c["row"][7] for c in fu if not c["exc"]), "%.1fs" % (time.time() - t0))

    wn = write_name_cases(vm, np.random.default_rng(52), 60)
    np.savez_compressed(os.path.join(OUT, "ref_written_names.npz"), name=np.array([c["name"] for c in wn]), stranded=np.array([c["stranded"] for c in wn]),
                        quals=np.array([c["quals"] for c in wn]), rev=np.array([c["rev"] for c in wn], dtype=np.int32),
                        five=np.array([c["five"] for c in wn], dtype=np.int32), read_id=np.array([c["rid"] for c in wn], dtype=np.int64),
                        kw=np.array([c["kw"] for c in wn]))
    print("getRecordForWriting", len(wn), "names")

    ex = exact_lookup_cases(vm, np.random.default_rng(8), 80)
    np.savez_compressed(os.path.join(OUT, "ref_exact_lookup.npz"), three_prime=np.array([c["tp"] for c in ex], dtype=np.int32),
                        whitelist=np.array([c["whitelist"] for c in ex], dtype=np.uint64), read=np.array([[r[0] for r in c["reads"]] for c in ex]),
                        adapterpos=np.array([[r[1] for r in c["reads"]] for c in ex], dtype=np.int32),
                        found=np.array([[r[2] for r in c["reads"]] for c in ex], dtype=np.int32),
                        count_keys=np.array([sorted(c["counts"]) + [0] * (80 - len(c["counts"])) for c in ex], dtype=np.uint64),
                        count_vals=np.array([[c["counts"][k] for k in sorted(c["counts"])] + [0] * (80 - len(c["counts"])) for c in ex], dtype=np.int64))
    print("pass-1 exact lookup", [(sum(r[2] == 1 for r in c["reads"]), sum(r[2] == -1 for r in c["reads"])) for c in ex], "(found, throwing) per geometry")

    tb = test_barcodes_cases(vm, np.random.default_rng(909), 24)
    save_test_barcodes(tb)
    print("testBarcodes", len(tb), "reads, with hits", sum(c["n_raw"] > 0 for c in tb), "with second", sum(c["n_distinct"] == 2 for c in tb), "%.1fs" % (time.time() - t0))

    gm = getmaxed_cases(vm)
    np.savez_compressed(os.path.join(OUT, "ref_getmaxed.npz"), rows=gm)
    print("getmaxED", gm.shape, "result histogram", np.bincount(gm[:, 5] + 1))

    dj = dojob_cases(vm, rng, 45)
    keys, koff = flat(dj, "keys")
    res = np.array([(i,) + r for i, c in enumerate(dj) for r in c["res"]], dtype=np.int64).reshape(-1, 8)
    np.savez_compressed(os.path.join(OUT, "ref_dojob.npz"), keys=keys, key_offsets=koff, w=np.array([c["w"] for c in dj], dtype=np.uint64),
                        ed=np.array([c["ed"] for c in dj], dtype=np.int32), mode=np.array([c["mode"] for c in dj], dtype=np.int32),
                        post=np.array([c["post"].ljust(5, "-") for c in dj]), off=np.array([c["off"] for c in dj], dtype=np.int32), res=res)
    print("doJob", len(dj), "cases,", len(res), "matches, %.1fs" % (time.time() - t0), vm.n_insn, "bytecodes")

    ac = assign_cases(vm, np.random.default_rng(777), 40)
    keys, koff = flat(ac, "keys")
    status = np.array([2 if isinstance(c["res"], str) else (0 if c["res"] is None else 1) for c in ac], dtype=np.int32)     # 0 unassigned, 1 assigned, 2 exception
    resrows = np.array([list(c["res"]) if isinstance(c["res"], tuple) else [0] * 7 for c in ac], dtype=np.uint64)
    cnt = np.array([(i, k, e, n) for i, c in enumerate(ac) for k, (tot, eds) in c["counts"].items() for e, n in eds.items()], dtype=np.int64).reshape(-1, 4)
    np.savez_compressed(os.path.join(OUT, "ref_assign.npz"), read=np.array([c["read"] for c in ac]), adapterpos=np.array([c["ap"] for c in ac], dtype=np.int32),
                        three_prime=np.array([c["tp"] for c in ac], dtype=np.int32), ed=np.array([c["ed"] for c in ac], dtype=np.int32),
                        pm=np.array([c["pm"] for c in ac], dtype=np.int32), keys=keys, key_offsets=koff, status=status, result=resrows, counts=cnt,
                        exc=np.array([c["res"] if isinstance(c["res"], str) else "" for c in ac]))
    print("assignBarcode", len(ac), "reads: unassigned/assigned/exception", np.bincount(status, minlength=3), "%.1fs" % (time.time() - t0), vm.n_insn, "bytecodes")

    gc = guided_cases(vm, rng, 36)
    keys, koff = flat(gc, "keys")
    ak, aoff = flat(gc, "allk")
    ek, eoff = flat(gc, "empk")
    res = np.array([(i,) + r for i, c in enumerate(gc) for r in c["res"]], dtype=np.int64).reshape(-1, 7)
    np.savez_compressed(os.path.join(OUT, "ref_guided.npz"), keys=keys, key_offsets=koff, all_keys=ak, all_offsets=aoff, empty_keys=ek, empty_offsets=eoff,
                        w=np.array([c["w"] for c in gc], dtype=np.uint64), L=np.array([c["L"] for c in gc], dtype=np.int32),
                        ed=np.array([c["ed"] for c in gc], dtype=np.int32), bc=np.array([c["bc"] for c in gc], dtype=np.int32),
                        bail=np.array([c["bail"] for c in gc], dtype=np.int32), off=np.array([c["off"] for c in gc], dtype=np.int32),
                        post=np.array([c["post"].ljust(10, "-") for c in gc]), res=res,
                        flag_gene=np.int64(512), flag_all=np.int64(4), flag_empty=np.int64(8))     # BarcodeFindingFlag.flagValue read from the enum
    print("guided", len(gc), "cases,", len(res), "list entries, %.1fs" % (time.time() - t0), vm.n_insn, "bytecodes")


if __name__ == "__main__":
    main()
