"""The oracle against vectors produced by the REFERENCE'S OWN BYTECODE (tests/golden/ref_*.npz): the class files of
/root/reference/Jar were executed by oracle/minijvm.py (a JVM-subset interpreter, there is no JDK in the image) with
oracle/make_ref_vectors.py; inputs and outputs are frozen here, nothing below needs the jars.  These vectors pin the 2-bit primitives,
BarcodeMatchTester.doJob (second-pass and collision-tester settings), the Illumina-guided engine with both checkMatchWithTestSets
flavours and the thresholded Levenshtein distance of the UMI matrix."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# narrow sets (oracle/make_ref_vectors.py) + wide sets with other seeds (oracle/make_ref_guided_wide.py), same layout
FIND_UMI_FILES = ["ref_find_umi.npz", "ref_find_umi_wide.npz"]
TEST_BARCODES_FILES = ["ref_test_barcodes.npz", "ref_test_barcodes_wide.npz"]
GUIDED_FILES = ["ref_guided.npz", "ref_guided_wide.npz"]
DOJOB_FILES = ["ref_dojob.npz", "ref_dojob_wide.npz"]                  # wide: oracle/make_ref_wide2.py
UMI_PAIR_FILES = ["ref_umi_pairs.npz", "ref_umi_pairs_wide.npz"]
HIER_FILES = ["ref_hier.npz", "ref_hier_wide.npz"]                     # wide: oracle/make_ref_hier_wide.py
MYCLUST_FILES = ["ref_myclust.npz", "ref_myclust_wide.npz"]            # wide: oracle/make_ref_myclust_wide.py
CLUSTER_LOCAL_FILES = ["ref_cluster_local.npz", "ref_cluster_local_wide.npz"]      # wide: oracle/make_ref_cluster_wide.py
M64 = (1 << 64) - 1


def test_primitives_match_reference_bytecode(orc):
    z = np.load(os.path.join(GOLDEN, "ref_primitives.npz"))
    L = orc.lib()
    for s, n, h, rc in zip(z["seq"], z["seq_len"], z["seq_hash"], z["seq_revcomp"]):
        s = str(s)[:int(n)]
        bad = C.c_int(0)
        assert L.orc_pack2bit(s.encode(), int(n), C.byref(bad)) == int(h) == pyref.pack(s)          # getLongHashForSeq incl. the N sign extension
        assert L.orc_revcomp2bit(int(h), int(n)) == int(rc) == pyref.revcomp2(int(h), int(n))        # reverseComplement()
    out = (C.c_uint64 * 4)()
    for row in z["rows"]:
        n, h, p = int(row[0]), int(row[1]), int(row[2])
        L.orc_replace_deg(h, out, p, n)
        assert list(out) == [int(x) for x in row[4:8]] == [x & M64 for x in pyref.replace_deg(h, p, n)]
        if p < n - 1:
            L.orc_insert_deg(h, out, p, n)
            assert list(out) == [int(x) for x in row[8:12]] == [x & M64 for x in pyref.insert_deg(h, p, n)]      # incl. the p = L-2 shift quirk
        for k, c4 in enumerate((1, 2, 4, 8, 15)):
            assert L.orc_delete_byte(h, c4, p, n) == int(row[12 + k]) == pyref.delete_byte(h, c4, p, n) & M64
    assert len(z["rows"]) == 1680


@pytest.mark.parametrize("fname", ["ref_levenshtein.npz", "ref_levenshtein_wide.npz"])      # wide: oracle/make_ref_small_wide.py
def test_limited_compare_matches_reference_bytecode(orc, fname):
    z = np.load(os.path.join(GOLDEN, fname))["rows"]
    L = orc.lib()
    seen = set()
    for row in z:
        a, b, d = row[:12].tobytes(), row[12:24].tobytes(), int(np.int8(row[24]))
        assert L.orc_limited_compare(a, 12, b, 12, 4) == d
        seen.add(d)
    assert seen == {-1, 0, 1, 2, 3, 4}


@pytest.mark.parametrize("fname", DOJOB_FILES)
def test_dojob_matches_reference_bytecode(orc, fname):
    """BarcodeMatchTester.doJob run by the reference's own class files: every OneMatch (readSeq, matchingBC, ED, counters, offset) in
    discovery order, for the second-pass settings (post sequence, doNextLevelIfMatchFound) and the collision tester's
    (skipFullMatches, postSeq = null, !doNext)"""
    z = np.load(os.path.join(GOLDEN, fname))
    n_cases, n_match = len(z["w"]), 0
    for i in range(n_cases):
        keys = z["keys"][z["key_offsets"][i]:z["key_offsets"][i + 1]]
        mode, ed, off, w = int(z["mode"][i]), int(z["ed"][i]), int(z["off"][i]), int(z["w"][i])
        post = None if mode == 2 else [pyref.ENCODE[ord(c)] for c in str(z["post"][i])[:5]]
        got, _ = orc.match_tester(orc.BarcodeSet(keys), w, 16, ed, skip_full=(mode == 2), allow_indels=True, post4=post, do_next=(mode != 2), offset=off)
        exp = z["res"][z["res"][:, 0] == i][:, 1:]
        got_rows = [(m["read_seq"], m["bc"], m["ed"], m["n_sub"], m["n_ins"], m["n_del"], m["offset"]) for m in got]
        assert got_rows == [tuple(int(x) & M64 if k < 2 else int(x) for k, x in enumerate(r)) for r in exp], (i, mode, ed, got_rows, exp)
        n_match += len(exp)
    assert (n_cases, n_match >= 60) == (45, True) if "wide" not in fname else (n_cases >= 240 and n_match >= 300)
    assert {int(e) for e in z["res"][:, 3]} == {0, 1, 2}                        # hits at every ED level


@pytest.mark.parametrize("fname", GUIDED_FILES)
def test_guided_engine_matches_reference_bytecode(orc, fname):
    """UMInuc / BCnucTwoBitPerBaseEDtester.matchesSeqEditDistance run by the reference's own class files: the whole ArrayList in list order
    (duplicates, counters, startOffsetFromPredicted, findingErrorFlag incl. the GENE bit inherited by descendants), with and without bailout"""
    z = np.load(os.path.join(GOLDEN, fname))
    fg, fa, fe = int(z["flag_gene"]), int(z["flag_all"]), int(z["flag_empty"])
    n_entries = 0
    flags_seen = set()
    for i in range(len(z["w"])):
        sl = lambda k, o: z[k][z[o][i]:z[o][i + 1]]
        keys, allk, empk = sl("keys", "key_offsets"), sl("all_keys", "all_offsets"), sl("empty_keys", "empty_offsets")
        bc, L, ed, bail, off, w = bool(z["bc"][i]), int(z["L"][i]), int(z["ed"][i]), int(z["bail"][i]), int(z["off"][i]), int(z["w"][i])
        post4 = [pyref.ENCODE[ord(c)] for c in str(z["post"][i]).rstrip("-")]
        got, n = orc.guided_tester(orc.BarcodeSet(keys), w, L, ed, post4, bailout=bail, offset=off, bc_flavour=bc,
                                   all_set=orc.BarcodeSet(allk) if bc else None, all_ed=3, empty_set=orc.BarcodeSet(empk) if bc else None, empty_ed=2)
        exp = z["res"][z["res"][:, 0] == i][:, 1:]
        assert n == len(exp), (i, n, len(exp))
        for g, e in zip(got, exp):
            where = (orc.W_GENE if int(e[5]) & fg else 0) | (orc.W_ALL if int(e[5]) & fa else 0) | (orc.W_EMPTY if int(e[5]) & fe else 0)
            assert (int(g["seq"]), g["n_sub"], g["n_ins"], g["n_del"], g["offset"], g["where"]) == (int(e[0]) & M64, e[1], e[2], e[3], e[4], where), (i, g, e)
            flags_seen.add(where)
        # the independent Python restatement as well
        sets = dict(umis=set(int(k) for k in keys)) if not bc else dict(gene=set(int(k) for k in keys) or None, all_bcs=set(int(k) for k in allk), all_ed=3,
                                                                        empty=set(int(k) for k in empk), empty_ed=2)
        t = pyref.GuidedTester(ed, L, post4, bailout=None if bail < 0 else bail, bc_flavour=bc, **sets)
        py = [(n_.seq, n_.nSub, n_.nIns, n_.nDel, n_.offset, n_.flag) for n_ in t.run(w, off)]
        assert py == [(int(g["seq"]), g["n_sub"], g["n_ins"], g["n_del"], g["offset"], g["where"]) for g in got]
        n_entries += len(exp)
    assert n_entries >= 100
    assert {orc.W_GENE, orc.W_ALL, orc.W_EMPTY, orc.W_GENE | orc.W_ALL} <= flags_seen      # the inherited GENE bit occurs in the reference's own output


@pytest.mark.parametrize("fname", ["ref_best9.npz", "ref_best9_wide.npz"])
def test_best_of_nine_and_packing_match_reference_bytecode(orc, fname):
    """new ClusteringEditDistanceBase(eds) run by the reference's own class files (eds from its LevenshteinDistance.apply): the visiting
    order ZERO, PLUSONE, MINUSONE with strict '<', the BestEditDistance int and its transposed copy, the equality constant"""
    z = np.load(os.path.join(GOLDEN, fname))
    L = orc.lib()
    assert L.orc_umi_equality() == int(z["equality"])
    shifted = 0
    for row in z["rows"]:
        a, b = row[:14].astype(np.uint8).tobytes(), row[14:28].astype(np.uint8).tobytes()
        packed = L.orc_umi_best9(a, b, 12)
        assert packed == int(np.int32(row[28])), (row, hex(packed))
        assert L.orc_umi_transpose(packed) == int(np.int32(row[29]))
        shifted += (packed >> 24) != 0x12
    assert shifted > 20                                   # best distance found at a non-central window in a good share of the pairs


def test_assign_barcode_matches_reference_bytecode(orc):
    """Parser.assignBarcode run by the reference's own class files on whole reads (3' and 5' geometry, several offsets / ED levels, N around
    the window, windows that run off the read): assigned barcode, ed, ed_sec, bcStart, bcEnd, rank, the BarcodeCounts update, the unassigned
    and the throwing reads.  The HashSet<OneMatch> iteration order behind `matches.stream().sorted()` is the JDK algorithm as modelled by
    oracle/minijvm.JdkHashSet (there is no JDK here); everything else is the reference's bytecode."""
    z = np.load(os.path.join(GOLDEN, "ref_assign.npz"))
    n = len(z["read"])
    seen = np.zeros(3, dtype=int)
    for i in range(n):
        read = str(z["read"][i]).encode()
        keys = z["keys"][z["key_offsets"][i]:z["key_offsets"][i + 1]]
        rank = np.arange(1, len(keys) + 1, dtype=np.int32)
        ap, tp, ed, pm = int(z["adapterpos"][i]), bool(z["three_prime"][i]), int(z["ed"][i]), int(z["pm"][i])
        anchor = (ap - 16) - 1 if tp else ap                    # Parser.java:L206-L210 (1-based adapterpos -> 0-based window start)
        sl = np.frombuffer(read, dtype=np.uint8).reshape(1, -1).copy()
        res, _ = orc.assign_barcode_batch(orc.BarcodeSet(keys, rank), sl, np.array([anchor], dtype=np.int32), ed, pm, tp, slice_len=len(read))
        r = res[0]
        status = int(z["status"][i])
        seen[status] += 1
        if status == 2:
            assert r["flags"] & orc.F_EXCEPTION, (i, str(z["exc"][i]), r)
            continue
        assert not (r["flags"] & orc.F_EXCEPTION), (i, r)
        if status == 0:
            assert not (r["flags"] & orc.F_ASSIGNED), (i, r)
            continue
        bc, e, e2, start, end, rk, _flag = (int(x) for x in z["result"][i])
        assert r["flags"] & orc.F_ASSIGNED, (i, r)
        off = int(r["offset"])
        g_start = ap - 1 + off if tp else ap + 1 + off          # Parser.java:L274-L276
        g_end = g_start - 15 - (int(r["n_ins"]) - int(r["n_del"])) if tp else g_start + 15 + (int(r["n_ins"]) - int(r["n_del"]))   # L278-L279
        assert (int(r["bc"]), int(r["ed"]), int(r["ed_second"]), g_start, g_end, int(r["rank"])) == (bc, e, int(np.int32(e2)), start, end, rk), (i, r, z["result"][i])
        cnt = z["counts"][z["counts"][:, 0] == i]
        assert len(cnt) == 1 and int(cnt[0, 1]) == bc and int(cnt[0, 2]) == e and int(cnt[0, 3]) == 1      # BarcodeCounts.addCountForEd (L305-L311)
    assert seen[1] >= 10 and seen[0] >= 3 and seen[2] >= 1, seen


def _check_wide_row(i, r, status, row, ap, tp, what):
    """one record (oracle or GPU) against the BarcodeResult fields Parser.assignBarcode produced in the interpreter"""
    if status == 2:
        assert r["flags"] & 2, (what, i, r)
        return
    assert not (r["flags"] & 2), (what, i, r)
    if status == 0:
        assert not (r["flags"] & 1), (what, i, r)
        return
    bc, e, e2, start, end, rk, _flag = (int(x) for x in row)
    off = int(r["offset"])
    g_start = ap - 1 + off if tp else ap + 1 + off              # Parser.java:L274-L276
    d = int(r["n_ins"]) - int(r["n_del"])
    g_end = g_start - 15 - d if tp else g_start + 15 + d        # L278-L279
    assert r["flags"] & 1 and (int(r["bc"]), int(r["ed"]), int(r["ed_second"]), g_start, g_end, int(r["rank"])) == \
           (bc, e, int(np.int32(e2)), start, end, rk), (what, i, r, row)


def _wide_constructed(z):
    for t in range(len(z["c_read"])):
        keys = z["c_keys"][z["c_key_offsets"][t]:z["c_key_offsets"][t + 1]]
        yield t, str(z["c_read"][t]).encode(), keys, np.arange(1, len(keys) + 1, dtype=np.int32), int(z["c_ap"][t]), bool(z["c_tp"][t]), int(z["c_pm"][t])


def test_assign_barcode_wide_matches_reference_bytecode(orc):
    """>= 550 whole reads through the reference's own Parser.assignBarcode bytecode at --bcEditDistance 2 (oracle/make_ref_assign_wide.py):
    the first reads of bench.py's bc3m_ed2 workload against the part of the 3 M list the reference can probe for them (so the frozen
    results are its results on the full list), and constructed periodic reads whose merged HashSet<OneMatch> holds 9 ... 27 entries
    (same-bin chains, 16 -> 32 -> 64 resizes, +-1 ... +-4 windows)."""
    z = np.load(os.path.join(GOLDEN, "ref_assign_wide.npz"))
    n = len(z["anchor"])
    assert n >= 500 and len(z["c_read"]) >= 50
    res, _ = orc.assign_barcode_batch(orc.BarcodeSet(z["U"], z["U_rank"]), z["slices"], z["anchor"], 2, 2, True)
    for i in range(n):
        _check_wide_row(i, res[i], int(z["status"][i]), z["result"][i], int(z["anchor"][i]) + 17, True, "oracle/bench")
    assert np.bincount(z["status"], minlength=3)[1] > 0.5 * n
    many = 0
    for t, read, keys, rank, ap, tp, pm in _wide_constructed(z):
        anchor = (ap - 16) - 1 if tp else ap
        sl = np.frombuffer(read, dtype=np.uint8).reshape(1, -1).copy()
        bset = orc.BarcodeSet(keys, rank)
        r, _ = orc.assign_barcode_batch(bset, sl, np.array([anchor], dtype=np.int32), 2, pm, tp, slice_len=len(read))
        _check_wide_row(t, r[0], int(z["c_status"][t]), z["c_result"][t], ap, tp, "oracle/constructed")
        entries = 0                                              # merged OneMatch entries = hits per (window, ED level)
        for o in range(-pm, pm + 1):
            ws = anchor + o
            if ws < 4 or ws + 21 > len(read):
                continue
            w = read[ws:ws + 16]
            if tp:
                seq = orc.lib().orc_revcomp2bit(orc.lib().orc_pack2bit(w, 16, None), 16)
            else:
                seq = orc.lib().orc_pack2bit(w, 16, None)
            post = read[ws - 4:ws + 1][::-1].translate(bytes.maketrans(b"ACGT", b"TGCA")) if tp else read[ws + 16:ws + 21]
            p4 = np.array([orc.lib().orc_encode4bit(c) for c in post], dtype=np.uint8)
            m, _ = orc.match_tester(bset, seq, 16, 2, post4=p4, offset=o)
            entries += 0 if m is None else len(m)
        many += entries >= 9
    assert many >= 20, many                                      # the resize / chain paths of slr_decide are really exercised


# ------------------------------------------------------------------------------------------------------------- GPU vs the reference's bytecode
@pytest.mark.gpu
def test_gpu_assign_barcode_wide_matches_reference_bytecode(pkg, ctx):
    """the CUDA kernel through the C ABI against the wide vectors: once with the probe-able part of the list as the table, once with the
    FULL 3 M list (same records: the reference cannot probe anything else for these reads), and the constructed many-entry reads"""
    z = np.load(os.path.join(GOLDEN, "ref_assign_wide.npz"))
    n = len(z["anchor"])
    wl = pkg.synth_whitelist(3_000_000, 3_000_000)
    for keys, rank, what in ((z["U"], z["U_rank"], "gpu/U"), (wl, np.arange(1, len(wl) + 1, dtype=np.int32), "gpu/3M")):
        table = pkg.BarcodesMapForBCfinding(ctx, keys, rank)
        res = pkg.Parser(ctx, table, bcEditDistance=2, testPlusMinusPos=2, three_prime=True).assign_barcodes(z["slices"], z["anchor"])
        for i in range(n):
            _check_wide_row(i, res[i], int(z["status"][i]), z["result"][i], int(z["anchor"][i]) + 17, True, what)
        table.close()
    for t, read, keys, rank, ap, tp, pm in _wide_constructed(z):
        anchor = (ap - 16) - 1 if tp else ap
        start = min(max(anchor - 10, 0), max(len(read) - 32, 0))
        piece = read[start:start + 32]
        sl = np.zeros((1, 32), dtype=np.uint8)
        sl[0, :len(piece)] = np.frombuffer(piece, dtype=np.uint8)
        table = pkg.BarcodesMapForBCfinding(ctx, keys, rank)
        r = pkg.Parser(ctx, table, bcEditDistance=2, testPlusMinusPos=pm, three_prime=tp).assign_barcodes(
            sl, np.array([anchor - start], dtype=np.int32), lens=np.array([len(piece)], dtype=np.int32))[0]
        # a 32-byte slice cannot hold +-4 windows with their flanks: those reads are oracle-only (the C ABI takes slices of <= 32 bytes)
        lo, hi = (anchor - pm - 4, anchor + pm + 16) if tp else (anchor - pm, anchor + pm + 21)
        if lo - start < 0 or hi - start > len(piece):
            continue
        _check_wide_row(t, r, int(z["c_status"][t]), z["c_result"][t], ap, tp, "gpu/constructed")
        table.close()


@pytest.mark.gpu
def test_gpu_assign_barcode_matches_reference_bytecode(pkg, ctx):
    """the CUDA kernel through the C ABI, directly against Parser.assignBarcode as run from the reference's class files"""
    z = np.load(os.path.join(GOLDEN, "ref_assign.npz"))
    checked = 0
    for i in range(len(z["read"])):
        read = str(z["read"][i]).encode()
        keys = z["keys"][z["key_offsets"][i]:z["key_offsets"][i + 1]]
        rank = np.arange(1, len(keys) + 1, dtype=np.int32)
        ap, tp, ed, pm = int(z["adapterpos"][i]), bool(z["three_prime"][i]), int(z["ed"][i]), int(z["pm"][i])
        anchor = (ap - 16) - 1 if tp else ap
        start = min(max(anchor - 8, 0), max(len(read) - 32, 0))               # a 32-byte slice of the read around the windows
        sl = np.zeros((1, 32), dtype=np.uint8)
        piece = read[start:start + 32]
        sl[0, :len(piece)] = np.frombuffer(piece, dtype=np.uint8)
        table = pkg.BarcodesMapForBCfinding(ctx, keys, rank)
        r = pkg.Parser(ctx, table, bcEditDistance=ed, testPlusMinusPos=pm, three_prime=tp).assign_barcodes(
            sl, np.array([anchor - start], dtype=np.int32), lens=np.array([len(piece)], dtype=np.int32))[0]
        status = int(z["status"][i])
        if status == 2:
            assert r["flags"] & 2, (i, r)
        elif status == 0:
            assert not (r["flags"] & 3), (i, r)
        else:
            bc, e, e2, bstart, bend, rk, _ = (int(x) for x in z["result"][i])
            off = int(r["offset"])
            g_start = ap - 1 + off if tp else ap + 1 + off
            g_end = g_start - 15 - (int(r["n_ins"]) - int(r["n_del"])) if tp else g_start + 15 + (int(r["n_ins"]) - int(r["n_del"]))
            assert r["flags"] & 1 and (int(r["bc"]), int(r["ed"]), int(r["ed_second"]), g_start, g_end, int(r["rank"])) == \
                   (bc, e, int(np.int32(e2)), bstart, bend, rk), (i, r, z["result"][i])
            counts = table.counts()
            assert counts.sum() == 1 and counts[list(keys).index(bc), e] == 1
        checked += 1
    assert checked == 40


@pytest.mark.parametrize("fname", GUIDED_FILES)
def test_sim_guided_engine_matches_reference_bytecode(sim, orc, fname):
    """the guided kernel's per-lane code (CPU replay) against the reference's own matchesSeqEditDistance lists: raw list in list order"""
    import workloads
    z = np.load(os.path.join(GOLDEN, fname))
    fg, fa, fe = int(z["flag_gene"]), int(z["flag_all"]), int(z["flag_empty"])
    n_entries = 0
    for i in range(len(z["w"])):
        sl_ = lambda k, o: z[k][z[o][i]:z[o][i + 1]]
        keys, allk, empk = sl_("keys", "key_offsets"), sl_("all_keys", "all_offsets"), sl_("empty_keys", "empty_offsets")
        bc, L, ed, bail, w = bool(z["bc"][i]), int(z["L"][i]), int(z["ed"][i]), int(z["bail"][i]), int(z["w"][i])
        post = str(z["post"][i]).rstrip("-")
        s = (workloads.g_unpack(w, L) + post).ljust(32, "A").encode()
        res, raw = _sim_guided_one(sim, orc, keys, L, np.frombuffer(s, dtype=np.uint8).reshape(1, 32), np.array([0], dtype=np.int32), ed, 0, len(post),
                                   bail, 32, bc=bc, allk=allk if bc else None, empk=empk if bc else None)
        exp = z["res"][z["res"][:, 0] == i][:, 1:]
        assert not res[0]["flags"] and res[0]["n_raw"] == len(exp), (i, res[0], len(exp))
        for g, e in zip(raw[0], exp[:64]):
            where = (1 if int(e[5]) & fg else 0) | (2 if int(e[5]) & fa else 0) | (4 if int(e[5]) & fe else 0)
            assert (int(g["seq"]), g["n_sub"], g["n_ins"], g["n_del"], g["where"]) == (int(e[0]) & M64, e[1], e[2], e[3], where), (i, g, e)
        n_entries += len(exp)
    assert n_entries >= (400 if "wide" in fname else 100)


@pytest.mark.gpu
@pytest.mark.parametrize("fname", GUIDED_FILES)
def test_gpu_guided_engine_matches_reference_bytecode(pkg, ctx, fname):
    """the guided kernel's raw list (plusminus 0: one tester) against UMInuc / BCnucTwoBitPerBaseEDtester.matchesSeqEditDistance as run from the
    reference's class files; startOffsetFromPredicted is a label the caller passes and is not compared"""
    import workloads
    z = np.load(os.path.join(GOLDEN, fname))
    fg, fa, fe = int(z["flag_gene"]), int(z["flag_all"]), int(z["flag_empty"])
    for i in range(len(z["w"])):
        sl_ = lambda k, o: z[k][z[o][i]:z[o][i + 1]]
        keys, allk, empk = sl_("keys", "key_offsets"), sl_("all_keys", "all_offsets"), sl_("empty_keys", "empty_offsets")
        bc, L, ed, bail, w = bool(z["bc"][i]), int(z["L"][i]), int(z["ed"][i]), int(z["bail"][i]), int(z["w"][i])
        post = str(z["post"][i]).rstrip("-")
        s = (workloads.g_unpack(w, L) + post).ljust(32, "A").encode()
        sets = pkg.GuidedSets(ctx, keys, np.array([0, len(keys)], dtype=np.int64), L, bc_flavour=bc, all_keys=allk if bc else None, all_ed=3,
                              empty_keys=empk if bc else None, empty_ed=2)
        res, raw = sets.match(np.frombuffer(s, dtype=np.uint8).reshape(1, 32), np.array([0], dtype=np.int32), np.array([0], dtype=np.int32), ed, 0,
                              len(post), bailout=None if bail < 0 else bail, slice_len=32, raw_cap=64)
        exp = z["res"][z["res"][:, 0] == i][:, 1:]
        assert not res[0]["flags"] and res[0]["n_raw"] == len(exp), (i, res[0], len(exp))
        for g, e in zip(raw[0], exp[:64]):
            where = (1 if int(e[5]) & fg else 0) | (2 if int(e[5]) & fa else 0) | (4 if int(e[5]) & fe else 0)
            assert (int(g["seq"]), g["n_sub"], g["n_ins"], g["n_del"], g["where"]) == (int(e[0]) & M64, e[1], e[2], e[3], where), (i, g, e)


@pytest.mark.parametrize("fname", UMI_PAIR_FILES)
def test_calc_edit_distances_matches_reference_bytecode(orc, sim, fname):
    """calcEditDistances itself (ClusteringEditDistanceBase.lambda$static$7) run by the reference's class files: stranded mini-sequence (3' reads
    through getSeqRevComp), the three windows getSubSequence(bcEnd + 1 + i, 12), the equal-bytes shortcut, nine distances, best-of-9, packing.
    The S2 boundary hands the kernel the 14 codes stranded[bcEnd - 1, bcEnd + 13) — exactly what is checked here."""
    z = np.load(os.path.join(GOLDEN, fname))
    L = orc.lib()
    codes = lambda s, e: bytes(pyref.ENCODE[ord(c)] for c in s[e - 1:e + 13])
    nonzero = 0
    n = len(z["packed"])
    umis = np.zeros((2 * n, 16), dtype=np.uint8)
    for i, (s1, s2, e1, e2, packed) in enumerate(zip(z["s1"], z["s2"], z["end1"], z["end2"], z["packed"])):
        a, b = codes(str(s1), int(e1)), codes(str(s2), int(e2))
        assert len(a) == 14 and len(b) == 14
        assert L.orc_umi_best9(a, b, 12) == int(np.int32(packed)), (s1, s2, e1, e2, hex(int(packed)))
        nonzero += (int(packed) & 0xFFFFFF) != 0
        umis[2 * i, :14], umis[2 * i + 1, :14] = list(a), list(b)
    assert n == (2000 if "wide" in fname else 200) and nonzero > n // 2
    # the same pairs through the CPU replay of the UMI kernel's per-lane code (what the GPU computes): every pair is a job of two reads
    offs = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
    oo = np.arange(0, 4 * n + 1, 4, dtype=np.int64)
    got = np.zeros(4 * n, dtype=np.int32)
    sim.sim_umi_dist(umis.ctypes.data, 16, 12, offs.ctypes.data, n, got.ctypes.data, oo.ctypes.data)
    assert (got.reshape(n, 4)[:, 1] == z["packed"].astype(np.int32)).all()


@pytest.mark.gpu
def test_gpu_umi_distance_matches_reference_bytecode(pkg, ctx):
    gpu_umi_distance(pkg, ctx, UMI_PAIR_FILES[0])              # the wide set runs from tests/test_zz_late_gpu.py


def gpu_umi_distance(pkg, ctx, fname):
    """the UMI kernel through the C ABI against calcEditDistances as run from the reference's class files: every pair is a job of two reads"""
    z = np.load(os.path.join(GOLDEN, fname))
    n = len(z["packed"])
    umis = np.zeros((2 * n, 16), dtype=np.uint8)
    for i, (s1, s2, e1, e2) in enumerate(zip(z["s1"], z["s2"], z["end1"], z["end2"])):
        for k, (s, e) in enumerate(((str(s1), int(e1)), (str(s2), int(e2)))):
            umis[2 * i + k, :14] = [pyref.ENCODE[ord(c)] for c in s[e - 1:e + 13]]
    offs = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
    m, oo = pkg.generate_distance_matrices(ctx, umis, offs, 12)
    got = m.reshape(n, 4)[:, 1]                                  # cell (0, 1) of every 2 x 2 matrix
    assert (got == z["packed"].astype(np.int32)).all(), np.nonzero(got != z["packed"].astype(np.int32))[0][:5]


def _sim_guided_one(sim, orc, keys, L, sl, anchor, ed, pm, post_len, bail, slen, bc=False, allk=None, empk=None, raw_cap=64):
    """one read through the CPU replay of the guided kernel's per-lane code (tests/host_sim): what the GPU computes, without a GPU"""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    goff = np.array([0, len(keys)], dtype=np.int64)
    ak = None if allk is None else np.ascontiguousarray(allk, dtype=np.uint64)
    ek = None if empk is None else np.ascontiguousarray(empk, dtype=np.uint64)
    out = np.zeros(1, dtype=orc.GUIDED_RESULT)
    raw = np.zeros((1, raw_cap), dtype=orc.GUIDED_HIT)
    sl = np.ascontiguousarray(sl, dtype=np.uint8)
    anchor = np.ascontiguousarray(anchor, dtype=np.int32)
    gid = np.zeros(1, dtype=np.int32)
    edv = np.array([ed], dtype=np.int32)
    sim.sim_guided_batch(keys.ctypes.data, goff.ctypes.data, 1, None if ak is None else ak.ctypes.data, 0 if ak is None else len(ak), 3,
                         None if ek is None else ek.ctypes.data, 0 if ek is None else len(ek), 2, int(bc), L, pm, bail, post_len,
                         sl.ctypes.data, 32, slen, anchor.ctypes.data, gid.ctypes.data, edv.ctypes.data, 1, out.ctypes.data, raw.ctypes.data, raw_cap)
    return out, raw




def _find_umi_inputs(z, i):
    """the S4 boundary for read i: 32-byte stranded slice, anchor = window start of offset 0 (IlluminaUMIanalyzer.java:L102, L112-L124)"""
    s, bc_end, ed, pm = str(z["stranded"][i]), int(z["bc_end"][i]), int(z["ed"][i]), int(z["pm"][i])
    post_len = ed + pm + 2                                             # lengthPostUMIseq (L113)
    need = bc_end + pm + 12 + post_len
    if len(s) < need:
        s = s + "A" * 9                                                # "too short -> added some As" (L118-L124): the caller pads
    start = max(bc_end - pm - 1, 0)
    piece = s[start:start + 32].encode()
    sl = np.zeros((1, 32), dtype=np.uint8)
    sl[0, :len(piece)] = np.frombuffer(piece, dtype=np.uint8)
    return sl, np.array([bc_end - start], dtype=np.int32), ed, pm, post_len, min(len(piece), 32)


def _check_find_umi(z, i, r, pack):
    row = z["row"][i]
    found, flag, bseq, bsub, bins, bdel, boff, has_second = (int(x) for x in row[:8])
    if str(z["exc"][i]):
        assert r["flags"] & 1, (i, str(z["exc"][i]))
        return "exc"
    assert not r["flags"], (i, r)
    if not found and has_second:
        # MORE_THAN_ONE_MATCH: the list holds two distinct survivors whose NeedlemanWunsch alignments tie (nMismatchDiffBestvsSecondBest == 0), so
        # failedUMIfinding(flag) turns found off (java:L203, L208, L220).  The alignment step stays Java; the record must deliver both
        # survivors, and the best one's geometry is what the reference derived the read positions from (L205-L206)
        assert flag & 4 and flag & 2 and int(row[10]) == 0, (i, row)
        assert r["n_distinct"] == 2, (i, r, row)
        assert int(r["seq"][1]) == pack(str(z["second"][i])), (i, r, str(z["second"][i]))
        start = 500 - (1 + int(r["offset"][0]))
        assert (int(row[8]), int(row[9])) == (start, start - (12 - 1 + int(r["n_ins"][0]) - int(r["n_del"][0]))), (i, r, row)
        return "ambiguous"
    assert (r["n_distinct"] > 0) == bool(found), (i, r, row)            # otherwise found <=> the list is not empty (java:L220)
    if not found:
        assert flag & 2 or flag != 0                                    # UMI_NOT_FOUND was set
        return "none"
    assert (int(r["seq"][0]), r["n_sub"][0], r["n_ins"][0], r["n_del"][0], r["offset"][0]) == (bseq, bsub, bins, bdel, boff), (i, r, row)
    assert (r["n_distinct"] == 2) == bool(has_second), (i, r, row)
    if has_second:
        assert int(r["seq"][1]) == pack(str(z["second"][i])), (i, r, str(z["second"][i]))
    # read positions the host derives from the record (findUMI L205-L206, 3' reads): start = bcEndPosRead - (1 + offset), end = start - (11 + nIns - nDel)
    start = 500 - (1 + boff)
    assert (int(row[8]), int(row[9])) == (start, start - (12 - 1 + bins - bdel)), (i, row)
    return "second" if has_second else "one"


@pytest.mark.parametrize("fname", FIND_UMI_FILES)
def test_find_umi_matches_reference_bytecode(orc, fname):
    """IlluminaUMIanalyzer.findUMI as a whole, run by the reference's class files: offset loop and window / post geometry, the testers,
    getBestAndSecondBCorUMI (sorted().distinct(), Needleman alignments): found / not found, the best entry (sequence, counters, offset), whether a
    second-best exists and which sequence it is, and the read positions derived from the record"""
    import workloads
    z = np.load(os.path.join(GOLDEN, fname))
    assert len(z["ed"]) >= (150 if "wide" in fname else 40)
    kinds = set()
    for i in range(len(z["ed"])):
        sl, anchor, ed, pm, post_len, slen = _find_umi_inputs(z, i)
        umis = z["umis"][z["umi_offsets"][i]:z["umi_offsets"][i + 1]]
        res, _, _ = orc.guided_batch(umis, np.array([0, len(umis)], dtype=np.int64), sl, anchor, np.array([0], dtype=np.int32), ed, 12, pm, post_len,
                                     bailout=int(z["bail"][i]), slice_len=slen)
        kinds.add(_check_find_umi(z, i, res[0], workloads.g_pack))
    assert {"none", "one", "second"} <= kinds and ("ambiguous" in kinds or "wide" not in fname)


@pytest.mark.parametrize("fname", FIND_UMI_FILES)
def test_sim_find_umi_matches_reference_bytecode(sim, orc, fname):
    """the same vectors through the CPU replay of the guided kernel's per-lane code"""
    import workloads
    z = np.load(os.path.join(GOLDEN, fname))
    for i in range(len(z["ed"])):
        sl, anchor, ed, pm, post_len, slen = _find_umi_inputs(z, i)
        umis = z["umis"][z["umi_offsets"][i]:z["umi_offsets"][i + 1]]
        res, _ = _sim_guided_one(sim, orc, umis, 12, sl, anchor, ed, pm, post_len, int(z["bail"][i]), slen)
        _check_find_umi(z, i, res[0], workloads.g_pack)


@pytest.mark.gpu
@pytest.mark.parametrize("fname", FIND_UMI_FILES)
def test_gpu_find_umi_matches_reference_bytecode(pkg, ctx, fname):
    import workloads
    z = np.load(os.path.join(GOLDEN, fname))
    for i in range(len(z["ed"])):
        sl, anchor, ed, pm, post_len, slen = _find_umi_inputs(z, i)
        umis = z["umis"][z["umi_offsets"][i]:z["umi_offsets"][i + 1]]
        bail = int(z["bail"][i])
        res, _ = pkg.GuidedSets(ctx, umis, np.array([0, len(umis)], dtype=np.int64), 12).match(sl, anchor, np.array([0], dtype=np.int32), ed, pm, post_len,
                                                                                               bailout=None if bail < 0 else bail, slice_len=slen)
        _check_find_umi(z, i, res[0], workloads.g_pack)


def test_getmaxed_matches_reference_bytecode(pkg):
    """DynamicEditDistances.getmaxED run by the reference's class files on its own tables (bcMaxEditDistances.xml, umiMaxEditDistances.xml): every
    (length, error %) column x candidate counts x posplusminus x cap, incl. the NoSuchElementException cases — against slr_dyn_max_ed"""
    z = np.load(os.path.join(GOLDEN, "ref_getmaxed.npz"))["rows"]
    lib = pkg.gpu_lib()
    seen = set()
    for row in z:
        L, err, count, pm, cap, exp = (int(x) for x in row[:6])
        col = np.array([x for x in row[6:] if x >= 0], dtype=np.int64)
        got = lib.slr_dyn_max_ed(col.ctypes.data, len(col), count, pm, cap)
        assert got == exp, (row, got)
        seen.add(exp)
    assert seen == {-1, 0, 1, 2, 3, 4} and len(z) > 6000


def _exact_inputs(z, g):
    """32-byte slices of the reads around the predicted window (the S1 / pass-1 boundary), anchor relative to the slice"""
    tp = bool(z["three_prime"][g])
    reads = [str(r) for r in z["read"][g]]
    ap = z["adapterpos"][g].astype(np.int64)
    anchor = (ap - 16) - 1 if tp else ap                             # UsedCellBCListGenerator.java:L211-L215
    sl = np.zeros((len(reads), 32), dtype=np.uint8)
    rel, lens = np.zeros(len(reads), dtype=np.int32), np.zeros(len(reads), dtype=np.int32)
    for i, r in enumerate(reads):
        start = min(max(int(anchor[i]) - 8, 0), max(len(r) - 32, 0))
        piece = r[start:start + 32].encode()
        sl[i, :len(piece)] = np.frombuffer(piece, dtype=np.uint8)
        rel[i], lens[i] = int(anchor[i]) - start, len(piece)
    exp = {int(k): int(v) for k, v in zip(z["count_keys"][g], z["count_vals"][g]) if v > 0}
    return tp, sl, rel, lens, exp, int((z["found"][g] == -1).sum())


EXACT_FILES = ["ref_exact_lookup.npz", "ref_exact_lookup_wide.npz"]


@pytest.mark.parametrize("fname", EXACT_FILES)
def test_exact_lookup_matches_reference_bytecode(orc, fname):
    """pass 1: UsedCellBCListGenerator$Worker's per-read lambda run by the reference's class files (window at the predicted position, whitelist
    test, unfilteredUsedBarcodeMap counts, reads whose window leaves the read) against the oracle's exact lookup"""
    z = np.load(os.path.join(GOLDEN, fname))
    for g in range(len(z["three_prime"])):
        tp, sl, anchor, lens, exp, n_throw = _exact_inputs(z, g)
        wl = z["whitelist"][g]
        res = orc.exact_lookup_batch(orc.BarcodeSet(wl, np.arange(1, len(wl) + 1, dtype=np.int32)), sl, anchor, tp, lens=lens)
        got = {}
        for r in res[(res["flags"] & 1) != 0]:
            got[int(r["bc"])] = got.get(int(r["bc"]), 0) + 1
        assert got == exp and len(exp) > 20
        assert int(((res["flags"] & 2) != 0).sum()) == n_throw


@pytest.mark.gpu
def test_gpu_exact_lookup_matches_reference_bytecode(pkg, ctx):
    gpu_exact_lookup(pkg, ctx, EXACT_FILES[0])                     # the wide set runs from tests/test_zz_late_gpu.py


def gpu_exact_lookup(pkg, ctx, fname):
    z = np.load(os.path.join(GOLDEN, fname))
    for g in range(len(z["three_prime"])):
        tp, sl, anchor, lens, exp, n_throw = _exact_inputs(z, g)
        wl = z["whitelist"][g]
        table = pkg.BarcodesMapForBCfinding(ctx, wl, np.arange(1, len(wl) + 1, dtype=np.int32))
        res = pkg.UsedCellBCListGenerator(ctx, table, three_prime=tp).addFastqs(sl, anchor, lens=lens)
        counts = table.counts()[:, 0]                                # unfilteredUsedBarcodeMap = the ED-0 counters of the table
        assert {int(k): int(c) for k, c in zip(wl, counts) if c} == exp
        assert int(((res["flags"] & 2) != 0).sum()) == n_throw


def _test_barcodes_inputs(z, i):
    """the S4 boundary of the BC flavour: window of offset 0 starts right after the nbasesOfAdapterSeqInReadname (3) adapter bases of the stranded
    mini-sequence (IlluminaBarcodeAnalyzer.java:L285-L287), 10 post bases (L302)"""
    s, pm = str(z["stranded"][i]), int(z["pm"][i])
    start = max(3 - pm - 1, 0)
    piece = s[start:start + 32].encode()
    sl = np.zeros((1, 32), dtype=np.uint8)
    sl[0, :len(piece)] = np.frombuffer(piece, dtype=np.uint8)
    cut = lambda k, o: z[k][z[o][i]:z[o][i + 1]]
    return sl, np.array([3 - start], dtype=np.int32), min(len(piece), 32), cut("gene", "gene_offsets"), cut("all_keys", "all_offsets"), cut("empty_keys", "empty_offsets")


def _check_test_barcodes(z, i, r):
    fl = lambda w: (512 if w & 1 else 0) | (4 if w & 2 else 0) | (8 if w & 4 else 0)          # SLR_G_W_* -> BarcodeFindingFlag values
    if str(z["exc"][i]):
        assert r["flags"] & 1
        return "exc"
    assert not r["flags"] and int(r["n_raw"]) == int(z["n_raw"][i]), (i, r, z["n_raw"][i])
    assert int(r["min_err_gene"]) == int(z["min_err_gene"][i])                                # testBarcodes L312-L314: goodMatchFound iff <= 1
    if z["n_raw"][i] == 0:
        assert r["n_distinct"] == 0
        return "none"
    b = z["best"][i]
    assert (int(r["seq"][0]), r["n_sub"][0], r["n_ins"][0], r["n_del"][0], r["offset"][0], fl(int(r["where"][0]))) == tuple(int(x) for x in b), (i, r, b)
    assert int(r["n_distinct"]) == int(z["n_distinct"][i])
    if z["n_distinct"][i] == 2:
        assert int(r["seq"][1]) == int(z["second"][i][0])
    return "second" if z["n_distinct"][i] == 2 else "one"


@pytest.mark.parametrize("fname", TEST_BARCODES_FILES)
def test_test_barcodes_matches_reference_bytecode(orc, fname):
    """IlluminaBarcodeAnalyzer.testBarcodes (one gene) + getBestAndSecondBCorUMI(CELLBC) run by the reference's class files: BC-flavour offset loop
    and geometry, the three candidate lists with their ED limits, bailout, the comparator with scoreWhereFound (ascending!), distinct, the
    second-best match, the list size and the minimum error count of GENE entries"""
    z = np.load(os.path.join(GOLDEN, fname))
    assert len(z["ed"]) >= (80 if "wide" in fname else 20)
    kinds = set()
    for i in range(len(z["ed"])):
        sl, anchor, slen, gene, allk, empk = _test_barcodes_inputs(z, i)
        res, _, _ = orc.guided_batch(gene, np.array([0, len(gene)], dtype=np.int64), sl, anchor, np.array([0], dtype=np.int32), int(z["ed"][i]), 16, int(z["pm"][i]), 10,
                                     bailout=int(z["bail"][i]), bc_flavour=True, all_keys=allk, all_ed=3, empty_keys=empk, empty_ed=2, slice_len=slen)
        kinds.add(_check_test_barcodes(z, i, res[0]))
    assert {"none", "one", "second"} <= kinds or {"none", "second"} <= kinds


@pytest.mark.parametrize("fname", TEST_BARCODES_FILES)
def test_sim_test_barcodes_matches_reference_bytecode(sim, orc, fname):
    """the same vectors through the CPU replay of the guided kernel's per-lane code"""
    z = np.load(os.path.join(GOLDEN, fname))
    for i in range(len(z["ed"])):
        sl, anchor, slen, gene, allk, empk = _test_barcodes_inputs(z, i)
        res, _ = _sim_guided_one(sim, orc, gene, 16, sl, anchor, int(z["ed"][i]), int(z["pm"][i]), 10, int(z["bail"][i]), slen, bc=True, allk=allk, empk=empk)
        _check_test_barcodes(z, i, res[0])


@pytest.mark.gpu
@pytest.mark.parametrize("fname", TEST_BARCODES_FILES)
def test_gpu_test_barcodes_matches_reference_bytecode(pkg, ctx, fname):
    z = np.load(os.path.join(GOLDEN, fname))
    for i in range(len(z["ed"])):
        sl, anchor, slen, gene, allk, empk = _test_barcodes_inputs(z, i)
        bail = int(z["bail"][i])
        sets = pkg.GuidedSets(ctx, gene, np.array([0, len(gene)], dtype=np.int64), 16, bc_flavour=True, all_keys=allk, all_ed=3, empty_keys=empk, empty_ed=2)
        res, _ = sets.match(sl, anchor, np.array([0], dtype=np.int32), int(z["ed"][i]), int(z["pm"][i]), 10, bailout=None if bail < 0 else bail, slice_len=slen)
        _check_test_barcodes(z, i, res[0])


# ---- ClusterOne_MyClustering.clusterLocal (ClusterOne_MyClustering.java:L175-L219), SURVEY.md §8f-3 ---------------------
def _cluster_labels(rec, job_offsets):
    """cluster of a read = smallest member of the set of keys that chose the same entry (the grouping of L199 / L219)"""
    label = np.full(len(rec), -1, dtype=np.int32)
    for j in range(len(job_offsets) - 1):
        a, b = int(job_offsets[j]), int(job_offsets[j + 1])
        groups = {}
        for c in range(a, b):
            k = int(rec["best_key"][c])
            if k >= 0:
                groups.setdefault(k, []).append(c - a)
        for mem in groups.values():
            for x in mem:
                label[a + x] = min(mem)
    return label


def _check_cluster_records(z, rec):
    assert np.array_equal(_cluster_labels(rec, z["job_offsets"]), z["label"])
    has = np.array([(rec["best_key"][a:b] >= 0).any() for a, b in zip(z["job_offsets"][:-1], z["job_offsets"][1:])])
    assert np.array_equal(has.astype(np.int32), z["present"])                      # Optional.empty <=> no key at all


@pytest.mark.parametrize("fname", CLUSTER_LOCAL_FILES)
def test_cluster_local_matches_reference_bytecode(orc, fname):
    z = np.load(os.path.join(GOLDEN, fname))
    assert len(z["ed"]) >= 50 and (z["label"] >= 0).sum() > 500
    rec = np.zeros(len(z["member"]), dtype=orc.CLUSTER_REC)
    for j in range(len(z["ed"])):                                                  # ed differs per job: one oracle call each
        a, b = int(z["job_offsets"][j]), int(z["job_offsets"][j + 1])
        m = z["packed"][z["out_offsets"][j]:z["out_offsets"][j + 1]]
        rec[a:b] = orc.umi_cluster_batch(m, [0, b - a], [0, (b - a) ** 2], int(z["ed"][j]), z["member"][a:b], z["rank"][a:b])
    _check_cluster_records(z, rec)
    assert (rec["n_ties"] > 1).sum() > 50                                          # the order of the map did decide some of them
    # and the stream-by-stream Python restatement agrees as well
    for j in range(len(z["ed"])):
        a, b = int(z["job_offsets"][j]), int(z["job_offsets"][j + 1])
        n = b - a
        m = z["packed"][z["out_offsets"][j]:z["out_offsets"][j + 1]].reshape(n, n)
        got = pyref.cluster_local(m.tolist(), [i for i in range(n) if z["member"][a + i]], int(z["ed"][j]), pyref.fastutil_key_order)
        want = {}
        for i in range(n):
            if z["label"][a + i] >= 0:
                want.setdefault(int(z["label"][a + i]), set()).add(i)
        assert (got or set()) == {frozenset(v) for v in want.values()}


@pytest.mark.gpu
def test_gpu_cluster_local_matches_reference_bytecode(pkg, ctx):
    gpu_cluster_local(pkg, ctx, CLUSTER_LOCAL_FILES[0])            # the wide set runs from tests/test_zz_late_gpu.py


def gpu_cluster_local(pkg, ctx, fname):
    import torch
    z = np.load(os.path.join(GOLDEN, fname))
    m = len(z["member"])
    rec = np.zeros(m, dtype=pkg.UMI_CLUSTER_REC)
    d_m = torch.from_numpy(z["packed"]).cuda()
    d_mem, d_rank = torch.from_numpy(z["member"]).cuda(), torch.from_numpy(z["rank"]).cuda()
    d_cnt = torch.empty(m, dtype=torch.int32, device="cuda")
    d_rec = torch.empty(m * 4, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for j in range(len(z["ed"])):
        a, b = int(z["job_offsets"][j]), int(z["job_offsets"][j + 1])
        jo = torch.tensor([0, b - a], dtype=torch.int64, device="cuda")
        oo = torch.tensor([0, (b - a) ** 2], dtype=torch.int64, device="cuda")
        rc = pkg.gpu_lib().slr_umi_cluster_dev(ctx.h, d_m.data_ptr() + 4 * int(z["out_offsets"][j]), jo.data_ptr(), oo.data_ptr(), 1, b - a,
                                               int(z["ed"][j]), d_mem.data_ptr() + a, d_rank.data_ptr() + 4 * a, d_cnt.data_ptr() + 4 * a,
                                               d_rec.data_ptr() + 16 * a, st)
        assert rc == 0
        torch.cuda.synchronize()
    rec = d_rec.cpu().numpy().view(pkg.UMI_CLUSTER_REC).reshape(m)
    _check_cluster_records(z, rec)


# ------------------------------------------------------------------------------------------------------------- ClusterOneHierarchical
def _check_hier(z, rec_of_job):
    """records (oracle or GPU) against what ClusterOneHierarchical.call wrote through OneNanoporeResult.setAttribute in the interpreter:
    U8 = the centre read and the cluster's mean shift, U1, U2 (absent = -1), the PREDICTED_POS flag, the SKIPPED_HIGHCOMPLEXITY flag value,
    nUMIfoundClustering"""
    off = z["job_offsets"]
    stats = np.zeros(4, dtype=int)
    for j in range(len(off) - 1):
        a, b = int(off[j]), int(off[j + 1])
        rec = rec_of_job(j)
        assert len(rec) == b - a
        if rec["flags"][0] & 4:                               # the JVM's own result depends on identity hash codes: canonical order only
            stats[3] += 1
        for i in range(b - a):
            r = rec[i]
            want_assigned = int(z["assigned"][a + i])
            assert int(r["flags"] & 1) == want_assigned, (j, i, r)
            assert bool(r["flags"] & 2) == bool(z["flagval"][a + i] != 0 and not want_assigned) or want_assigned, (j, i, r, z["flagval"][a + i])
            if want_assigned:
                assert str(z["u8"][a + i]) == "UMI(%d,%d)" % (int(r["center"]), int(r[4])), (j, i, r, z["u8"][a + i])
                assert (int(r["u1"]), int(r["u2"]), int(r["pos2"])) == (int(z["u1"][a + i]), int(z["u2"][a + i]), int(z["pos2"][a + i])), (j, i, r)
                stats[0] += 1
            elif r["flags"] & 2:
                stats[1] += 1
        assert int(z["n_found"][j]) == int((rec["flags"] & 1).sum())
        stats[2] += 1
    return stats


@pytest.mark.parametrize("fname", HIER_FILES)
def test_cluster_one_hierarchical_matches_reference_bytecode(orc, fname):
    """ClusterOneHierarchical.call as a whole — LingPipe's CompleteLinkClusterer / SingleLinkClusterer / Dendrogram / BoundedPriorityQueue /
    ObjectToSet, DistanceMatrix, OneUmiCluster.setClusterCenter, ClusterOneBase.setSamflagsAndStatsForClustered — run from the reference's own
    class files (oracle/make_ref_hier.py) on jobs of 2 ... 100 reads: the C oracle reproduces every value the bytecode wrote"""
    z = np.load(os.path.join(GOLDEN, fname))
    off, oo = z["job_offsets"], z["out_offsets"]

    def rec_of_job(j):
        n = int(off[j + 1] - off[j])
        p = z["params"][j]
        return orc.umi_assign_batch(z["packed"][oo[j]:oo[j + 1]], np.array([0, n]), np.array([0, n * n]),
                                    orc.AssignParams(int(p[0]), int(p[1]), int(p[2]), int(p[3]), 100), z["qv01"][j:j + 1])
    stats = _check_hier(z, rec_of_job)
    assert stats[2] >= 200 and stats[0] > 1500 and stats[1] > 0, stats


@pytest.mark.gpu
def test_gpu_cluster_one_hierarchical_matches_reference_bytecode(pkg, ctx):
    stats = gpu_cluster_one_hierarchical(pkg, ctx, HIER_FILES[0])          # the wide set runs from tests/test_zz_late_gpu.py
    assert stats[2] >= 200 and stats[0] > 1500, stats


def gpu_cluster_one_hierarchical(pkg, ctx, fname):
    """the CUDA kernels through slr_umi_assign_dev against the same vectors (jobs grouped by parameter set)"""
    import ctypes as C
    import torch
    z = np.load(os.path.join(GOLDEN, fname))
    off, oo = z["job_offsets"], z["out_offsets"]
    recs = {}
    for prm in sorted({tuple(int(x) for x in p) for p in z["params"]}):
        js = [j for j in range(len(off) - 1) if tuple(int(x) for x in z["params"][j]) == prm]
        sizes = np.array([off[j + 1] - off[j] for j in js], dtype=np.int64)
        so = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        soo = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
        mats = np.concatenate([z["packed"][oo[j]:oo[j + 1]] for j in js]).astype(np.int32)
        qv = np.ascontiguousarray(z["qv01"][js])
        d_m, d_o, d_oo, d_q = (torch.from_numpy(x).cuda() for x in (mats, so, soo, qv))
        d_rec = torch.zeros((int(so[-1]), 16), dtype=torch.uint8, device="cuda")
        d_scr = torch.zeros(int(pkg.gpu_lib().slr_umi_assign_scratch_bytes(len(js))), dtype=torch.uint8, device="cuda")
        P = pkg.UmiAssignParams(prm[0], prm[1], prm[2], prm[3], 100)
        pkg._check(pkg.gpu_lib().slr_umi_assign_dev(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), len(js), int(so[-1]), C.byref(P),
                                                    d_q.data_ptr(), d_scr.data_ptr(), d_rec.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        got = d_rec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)
        for k, j in enumerate(js):
            recs[j] = got[so[k]:so[k + 1]]
    return _check_hier(z, lambda j: recs[j])


# ------------------------------------------------------------------------------------------------------------- ClusterOne_MyClustering
@pytest.mark.parametrize("fname", MYCLUST_FILES)
def test_cluster_one_myclustering_matches_reference_bytecode(orc, fname):
    """ClusterOne_MyClustering.call as a whole — clusterLocal on the full set, the depth rule, OneUmiCluster.setClusterCenter, the off-centre removal
    (removeEntries -> the fastutil iterator's backward-shift deletion -> re-centring), the second clusterLocal over the unclustered reads and
    setSamflagsAndStatsForClustered — run from the reference's own class files (oracle/make_ref_myclust.py, sequential streams) on 72 jobs of
    20 ... 330 reads: the C oracle reproduces every value the bytecode wrote"""
    z = np.load(os.path.join(GOLDEN, fname))
    off, oo = z["job_offsets"], z["out_offsets"]

    def rec_of_job(j):
        n = int(off[j + 1] - off[j])
        p = z["params"][j]
        rec = orc.umi_assign_batch(z["packed"][oo[j]:oo[j + 1]], np.array([0, n]), np.array([0, n * n]),
                                   orc.AssignParams(int(p[0]), int(p[1]), int(p[2]), int(p[3]), 0, 1), z["qv01"][j:j + 1])
        assert (rec["flags"] & 8).all()                       # ORC_UA_DEEP marks the records of ClusterOne_MyClustering
        return rec
    stats = _check_hier(z, rec_of_job)
    assert stats[2] >= 72 and stats[0] > 9000 and stats[1] > 0, stats


@pytest.mark.gpu
def test_gpu_cluster_one_myclustering_matches_reference_bytecode(pkg, ctx):
    stats = gpu_cluster_one_myclustering(pkg, ctx, MYCLUST_FILES[0])       # the wide set runs from tests/test_zz_late_gpu.py
    assert stats[2] >= 72 and stats[0] > 9000, stats


def gpu_cluster_one_myclustering(pkg, ctx, fname):
    """the large-job kernels (umi_assign_deep.cu) through slr_umi_assign_dev2 against the same vectors (max_hier = 0 sends every job there)"""
    import ctypes as C
    import torch
    z = np.load(os.path.join(GOLDEN, fname))
    off, oo = z["job_offsets"], z["out_offsets"]
    L = pkg.gpu_lib()
    recs = {}
    for prm in sorted({tuple(int(x) for x in p) for p in z["params"]}):
        js = [j for j in range(len(off) - 1) if tuple(int(x) for x in z["params"][j]) == prm]
        sizes = np.array([off[j + 1] - off[j] for j in js], dtype=np.int64)
        so = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        soo = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
        mats = np.concatenate([z["packed"][oo[j]:oo[j + 1]] for j in js]).astype(np.int32)
        qv = np.ascontiguousarray(z["qv01"][js])
        d_m, d_o, d_oo, d_q = (torch.from_numpy(x).cuda() for x in (mats, so, soo, qv))
        d_rec = torch.zeros((int(so[-1]), 16), dtype=torch.uint8, device="cuda")
        nbytes = int(L.slr_umi_assign_scratch_bytes(len(js))) + sum(int(L.slr_umi_assign_deep_job_bytes(int(n))) for n in sizes)
        d_scr = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        P = pkg.UmiAssignParams(prm[0], prm[1], prm[2], prm[3], 0, 1)
        pkg._check(L.slr_umi_assign_dev2(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), len(js), int(so[-1]), C.byref(P),
                                         d_q.data_ptr(), d_scr.data_ptr(), nbytes, d_rec.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        got = d_rec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)
        for k, j in enumerate(js):
            recs[j] = got[so[k]:so[k + 1]]
    return _check_hier(z, lambda j: recs[j])


def test_oversized_group_split_matches_reference_bytecode(pkg):
    """the caller-side split of a huge (cell, region) group (UmiClustering.lambda$cluster$7 + ListUtils.partition, run from the class files by
    oracle/make_ref_split.py) against the host mirror split_oversized_group"""
    z = np.load(os.path.join(GOLDEN, "ref_split.npz"))
    o = 0
    for ram, n, k in zip(z["ram"], z["n"], z["n_parts"]):
        assert pkg.split_oversized_group(int(n), int(ram)) == [int(x) for x in z["parts"][o:o + k]], (ram, n)
        o += int(k)
    assert len(z["n"]) >= 80 and int(z["n_parts"].max()) >= 10
