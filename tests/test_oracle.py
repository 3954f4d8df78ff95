"""CPU tests of the oracle (oracle/slr_oracle.c): reference README vectors, primitive known answers restated from the
bytecode, cross-check against the independent Python restatement (oracle/pyref.py), frozen golden vectors."""
import glob
import os
import random

import numpy as np
import pytest

import workloads
from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B = "AGCT"


def dec(x, L=16):
    return "".join(B[(int(x) >> (2 * (L - 1 - i))) & 3] for i in range(L))


# ---- the only input -> output pairs the reference publishes: /root/reference/README.md:400 and :452 -------------
README_VECTORS = [
    ("AAAAAAAAAAAATGGCGTGTATTGTCTTGGCACGATCGGAAGA", "TCCGATCGTGCCAAGA", 0, 619, 618, 603),
    ("AAAAAAAAAAACAAACCAAGTAACCAACCCAACCTCACTCAGA", "GAGTGAGGTTGGGTAG", 1, 1327, 1326, 1311),
]


@pytest.mark.parametrize("x,bc,ed,ae,bc_start,bc_end", README_VECTORS)
def test_readme_read_name_examples(orc, x, bc, ed, ae, bc_start, bc_end):
    rng = random.Random(3)
    wl = np.array([pyref.pack(bc)] + [rng.getrandbits(32) for _ in range(5000)], dtype=np.uint64)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    sl = np.frombuffer(x.encode(), dtype=np.uint8)[None, :].copy()
    # X= ends with 3 adapter bases, so the offset-0 window is X[-19:-3]
    res, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), sl, np.array([len(x) - 19], dtype=np.int32), 1)
    r = res[0]
    assert r["flags"] == orc.F_ASSIGNED and dec(r["bc"]) == bc and r["ed"] == ed and r["ed_second"] == orc.INT_MAX
    assert r["offset"] == 0 and r["rank"] == 1
    # bcStart / bcEnd as in the read name (Parser.java:L273-L280)
    start = ae - 1 + int(r["offset"])
    end = start - 15 - (int(r["n_ins"]) - int(r["n_del"]))
    assert (start, end) == (bc_start, bc_end)


# ---- 2-bit primitives (T!...NucleicAcidTwoBitPerBase) -------------------------------------------------------------
def test_pack_and_revcomp(orc):
    L = orc.lib()
    assert L.orc_pack2bit(b"AGCT", 4, None) == 0b00011011
    assert L.orc_pack2bit(b"agct", 4, None) == 0b00011011
    assert L.orc_revcomp2bit(pyref.pack("AAGC"), 4) == pyref.pack("GCTT")
    # a non-ACGT char ORs a sign-extended (byte)-2: high garbage, and every earlier digit becomes T, that digit C
    h = L.orc_pack2bit(b"ACNGT", 5, None)
    assert h == pyref.pack("ACNGT") and h >> 32 != 0
    assert (h & 0x3FF) == pyref.pack("TTCGT")
    # reverseComplement only reads the low 2L bits -> a 3' window with an N is a clean (wrong) barcode
    assert L.orc_revcomp2bit(h, 5) == pyref.pack("ACGAA")


def test_insert_shift_overflow_bug(orc):
    """getLongHashInsertByteDeg at pos = L-2: `>>> 64` is `>>> 0`, the old last base lands in bits 62-63."""
    import ctypes as C
    L = orc.lib()
    out = (C.c_uint64 * 4)()
    seq = pyref.pack("ACGTACGTACGTACGT")
    L.orc_insert_deg(seq, out, 14, 16)
    assert list(out) == pyref.insert_deg(seq, 14, 16)
    for v in range(4):
        assert out[v] >> 62 == 3                       # old last base T
        assert out[v] & 0xFFFFFFFF == (seq & ~3) | v   # low 32 bits: w[0..14] + v
    L.orc_insert_deg(pyref.pack("ACGTACGTACGTACGA"), out, 14, 16)
    assert all(o >> 32 == 0 for o in out)              # last base A: clean
    L.orc_insert_deg(seq, out, 3, 16)
    assert [dec(o) for o in out] == ["ACGTAACGTACGTACG", "ACGTGACGTACGTACG", "ACGTCACGTACGTACG", "ACGTTACGTACGTACG"]


def test_delete_and_replace(orc):
    import ctypes as C
    L = orc.lib()
    seq = pyref.pack("ACGTACGTACGTACGT")
    assert dec(L.orc_delete_byte(seq, 2, 0, 16)) == "CGTACGTACGTACGTG"      # code 2 = G appended
    assert dec(L.orc_delete_byte(seq, 15, 5, 16)) == "ACGTAGTACGTACGTA"     # N (15) appends A
    assert dec(L.orc_delete_byte(seq, 8, 14, 16)) == "ACGTACGTACGTACTT"
    out = (C.c_uint64 * 4)()
    L.orc_replace_deg(seq, out, 15, 16)
    assert [dec(o)[-1] for o in out] == list("AGCT")
    for pos in range(16):
        for code in (0, 1, 2, 4, 8, 15):
            assert L.orc_delete_byte(seq, code, pos, 16) == pyref.delete_byte(seq, code, pos, 16)
        L.orc_replace_deg(seq, out, pos, 16)
        assert list(out) == pyref.replace_deg(seq, pos, 16)
        if pos < 15:
            L.orc_insert_deg(seq, out, pos, 16)
            assert list(out) == pyref.insert_deg(seq, pos, 16)


# ---- BarcodeMatchTester.doJob ---------------------------------------------------------------------------------------
def test_probe_counts(orc):
    """App. A.5 of SURVEY.md: 124 probes per window at ED 1 (1 + 48 sub + 60 ins + 15 del)."""
    wl = np.array([1, 2, 3], dtype=np.uint64)
    m, probes = orc.match_tester(orc.BarcodeSet(wl), pyref.pack("ACGTTGCAGGTCAATC"), 16, 1, post4=[1, 2, 4, 8, 1])
    assert m == [] and probes == 124
    _, probes0 = orc.match_tester(orc.BarcodeSet(wl), pyref.pack("ACGTTGCAGGTCAATC"), 16, 0, post4=[1, 2, 4, 8, 1])
    assert probes0 == 1


def test_first_hit_wins_and_order(orc):
    """Two barcodes at ED 1: only the first in traversal order (position asc; SUB A,G,C,T; INS; DEL) is kept."""
    w = "ACGTTGCAGGTCAATC"
    a = "ACGTTGCAGGTCAATG"     # sub at position 15
    b = "AGGTTGCAGGTCAATC"     # sub at position 1 -> discovered first
    wl = np.array([pyref.pack(a), pyref.pack(b)], dtype=np.uint64)
    m, _ = orc.match_tester(orc.BarcodeSet(wl), pyref.pack(w), 16, 1, post4=[1, 1, 1, 1, 1])
    assert len(m) == 1 and dec(m[0]["bc"]) == b and m[0]["ed"] == 1 and m[0]["n_sub"] == 1


def test_same_barcode_in_two_ed_slots(orc):
    """App. B-2: a barcode reached at ED 2 before its ED-1 node is expanded occupies both slots."""
    w = "ACGTTGCAGGTCAATC"
    t = "ACGTTGCAGGTCACTC"     # ED 1 via sub at position 13; reached earlier at ED 2 (sub 0 -> X, then ... no: via ins+del paths)
    wl = np.array([pyref.pack(t)], dtype=np.uint64)
    m, _ = orc.match_tester(orc.BarcodeSet(wl), pyref.pack(w), 16, 2, post4=[1, 1, 1, 1, 1])
    p = pyref.BarcodeMatchTester(pyref.pack(w), 16, 2, False, True, {pyref.pack(t)}, 0, [1, 1, 1, 1, 1], True).doJob()
    assert sorted((x["ed"], x["bc"]) for x in m) == sorted((e.ed, e.bc) for e in p)
    assert 1 in [x["ed"] for x in m]


def test_last_position_substitution_quirk(orc):
    """Consequence of the insert shift bug + the (int) visited set (App. B-3/B-4): with ed >= 2 the level-1 node
    INS(pos 14, v) has low 32 bits == SUB(pos 15, v) but garbage in bits 62-63, so it never matches, yet it marks the
    value as tested -> the substitution at the last position is skipped at ED 1 and only found at ED 2."""
    w = "ACGTTGCAGGTCAATC"
    t = "ACGTTGCAGGTCAATG"
    wl = np.array([pyref.pack(t)], dtype=np.uint64)
    m1, _ = orc.match_tester(orc.BarcodeSet(wl), pyref.pack(w), 16, 1, post4=[1, 1, 1, 1, 1])
    assert [x["ed"] for x in m1] == [1]                    # no visited set at ed = 1
    m2, _ = orc.match_tester(orc.BarcodeSet(wl), pyref.pack(w), 16, 2, post4=[1, 1, 1, 1, 1])
    assert [x["ed"] for x in m2] == [2]
    wa = "ACGTTGCAGGTCAATA"                                # last base A: the inserted mutant is clean and hits at ED 1 as an indel
    m3, _ = orc.match_tester(orc.BarcodeSet(wl), pyref.pack(wa), 16, 2, post4=[1, 1, 1, 1, 1])
    e1 = [x for x in m3 if x["ed"] == 1]
    assert len(e1) == 1 and e1[0]["n_del"] == 1 and e1[0]["n_sub"] == 0


@pytest.mark.parametrize("three_prime", [True, False])
@pytest.mark.parametrize("ed", [0, 1, 2])
def test_oracle_vs_pyref(orc, three_prime, ed):
    """Same bytecode, two restatements (C with explicit stack / Python object-for-object) must agree on every field."""
    for seed, skew in ((1, False), (2, True)):
        reads, slices, anchors, wl = workloads.adversarial(seed * 7 + ed, three_prime, 30, skew=skew)
        rank = np.arange(1, len(wl) + 1, dtype=np.int32)
        ranks = {int(k): int(r) for k, r in zip(wl, rank)}
        res, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, three_prime)
        S = set(int(k) for k in wl)
        for i, r in enumerate(reads):
            ap = (8 + 16 + 1) if three_prime else 8
            o = res[i]
            try:
                p = pyref.assign_barcode(r, ap, S, ranks, ed, 2, three_prime)
            except pyref.JavaException:
                assert o["flags"] & orc.F_EXCEPTION, r
                continue
            exp = (p["bc"], p["ed"], p["ed_second"], p["offset"], p["n_ins"], p["n_del"], p["n_sub"], p["rank"], int(p["assigned"]))
            got = tuple(int(o[f]) for f in ("bc", "ed", "ed_second", "offset", "n_ins", "n_del", "n_sub", "rank", "flags"))
            assert exp == got, r


def test_exceptions(orc):
    wl = np.array([1, 2, 3], dtype=np.uint64)
    bs = orc.BarcodeSet(wl)
    s = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)[None, :].copy()
    ok, _ = orc.assign_barcode_batch(bs, s, np.array([8], dtype=np.int32), 1, 2, True)
    assert ok[0]["flags"] == 0 and ok[0]["ed"] == -1
    for anchor in (5, 15):                                  # 3': needs [anchor-6, anchor+18)
        r, _ = orc.assign_barcode_batch(bs, s, np.array([anchor], dtype=np.int32), 1, 2, True)
        assert r[0]["flags"] == orc.F_EXCEPTION
    for anchor in (1, 10):                                  # 5': needs [anchor-2, anchor+23)
        r, _ = orc.assign_barcode_batch(bs, s, np.array([anchor], dtype=np.int32), 1, 2, False)
        assert r[0]["flags"] == orc.F_EXCEPTION
    bad = s.copy()
    bad[0, 5] = ord("X")                                    # inside a 3' post sequence: ONEBYTE_REVERSECOMP_MATRIX[-1]
    r, _ = orc.assign_barcode_batch(bs, bad, np.array([8], dtype=np.int32), 1, 2, True)
    assert r[0]["flags"] == orc.F_EXCEPTION
    bad = s.copy()
    bad[0, 12] = ord("X")                                   # inside the window only: garbage hash, no exception
    r, _ = orc.assign_barcode_batch(bs, bad, np.array([8], dtype=np.int32), 1, 0, True)
    assert r[0]["flags"] == 0


def test_three_prime_n_window_can_match(orc):
    """3' mode: the N garbage is dropped by reverseComplement, the window becomes revcomp(T..TC + rest) and may match."""
    read = "ACGTACGTAGGNCATTGACCGTAAGGCTACGT"
    win = read[8:24]
    j = win.index("N")
    fake = "T" * j + "C" + win[j + 1:]
    bc = workloads.rcs(fake)
    wl = np.array([pyref.pack(bc), 12345], dtype=np.uint64)
    s = np.frombuffer(read.encode(), dtype=np.uint8)[None, :].copy()
    r, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), s, np.array([8], dtype=np.int32), 0, 0, True)
    assert r[0]["flags"] == orc.F_ASSIGNED and dec(r[0]["bc"]) == bc and r[0]["ed"] == 0
    r5, _ = orc.assign_barcode_batch(orc.BarcodeSet(np.array([pyref.pack(fake) & 0xFFFFFFFF], dtype=np.uint64)), s,
                                     np.array([8], dtype=np.int32), 0, 0, False)
    assert r5[0]["flags"] == 0 and r5[0]["ed"] == -1       # 5': bits >= 32 stay set, never matches


def test_hashset_order_breaks_ties(orc):
    """App. B-7: the same barcode best at two non-zero offsets with equal ED: HashSet iteration order decides the offset."""
    rng = random.Random(9)
    seen = set()
    for _ in range(300):
        core = "".join(rng.choice("ACGT") for _ in range(16))
        read = "".join(rng.choice("ACGT") for _ in range(8)) + core + "".join(rng.choice("ACGT") for _ in range(8))
        # barcode = revcomp of the window at offset -1 with its first base substituted: also ED 1 from offset -2 via an indel? not
        # guaranteed; just record which offset wins and compare with pyref
        bc_read = read[7:23]
        b = workloads.rcs(bc_read)
        b = b[:5] + ("A" if b[5] != "A" else "C") + b[6:]
        wl = np.array([pyref.pack(b)], dtype=np.uint64)
        s = np.frombuffer(read.encode(), dtype=np.uint8)[None, :].copy()
        r, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), s, np.array([8], dtype=np.int32), 2, 2, True)
        p = pyref.assign_barcode(read, 8 + 17, {int(wl[0])}, None, 2, 2, True)
        assert int(r[0]["offset"]) == p["offset"] and int(r[0]["ed"]) == p["ed"] and bool(r[0]["flags"] & 1) == p["assigned"]
        seen.add(int(r[0]["offset"]))
    assert -1 in seen


def test_golden_vectors(orc):
    files = sorted(glob.glob(os.path.join(GOLDEN, "bc_*.npz")))
    assert len(files) >= 7
    for f in files:
        g = np.load(f)
        res, probes = orc.assign_barcode_batch(orc.BarcodeSet(g["whitelist"], g["rank"]), g["slices"], g["anchor"], int(g["ed"]), 2,
                                               bool(g["three_prime"]))
        assert (res == g["result"]).all(), f
        assert probes == int(g["probes"]), f


# ---- UMI --------------------------------------------------------------------------------------------------------------
def lev(a, b):
    p = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        d = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            d[j] = min(d[j - 1] + 1, p[j] + 1, p[j - 1] + (a[i - 1] != b[j - 1]))
        p = d
    return p[len(b)]


def test_limited_compare_is_capped_levenshtein(orc):
    rng = random.Random(4)
    L = orc.lib()
    for _ in range(3000):
        n = rng.choice((10, 12))
        a = bytes(rng.choice((1, 2, 4, 8, 15)) for _ in range(n))
        b = bytearray(a)
        for _ in range(rng.randrange(0, 7)):
            b[rng.randrange(n)] = rng.choice((1, 2, 4, 8))
        if rng.random() < 0.5:
            k = rng.randrange(1, 4)
            b = b[k:] + bytes(rng.choice((1, 2, 4, 8)) for _ in range(k))
        b = bytes(b)
        d = lev(a, b)
        got = L.orc_limited_compare(a, n, b, n, 4)
        assert got == (d if d <= 4 else -1)
        assert got == pyref.limited_compare(list(a), list(b), 4)


def test_umi_best9_and_matrix(orc):
    L = orc.lib()
    umis, offs = workloads.umi_jobs(5, 12, n_jobs=12, max_n=12)
    m, oo = orc.umi_matrix_batch(umis, offs, 12)
    for j in range(len(offs) - 1):
        n = offs[j + 1] - offs[j]
        mat = m[oo[j]:oo[j + 1]].reshape(n, n)
        for i in range(n):
            assert mat[i, i] == L.orc_umi_equality() == (0 | (0x08000000 << 1) | (0x01000000 << 1))
            for v in range(i + 1, n):
                a, b = umis[offs[j] + i, :14], umis[offs[j] + v, :14]
                assert mat[i, v] == pyref.umi_best9(list(a), list(b), 12)
                assert mat[v, i] == L.orc_umi_transpose(int(mat[i, v]))
                assert (mat[v, i] & 0xFFFFFF) == (mat[i, v] & 0xFFFFFF) <= 5


def test_umi_visit_order_ties(orc):
    """strict '<' in the order ZERO, PLUSONE, MINUSONE: identical reads give (0, ZERO, ZERO); a tie between shifts keeps the first."""
    a = bytes([1, 2, 4, 8, 1, 2, 4, 8, 1, 2, 4, 8, 1, 2])
    e = orc.lib().orc_umi_best9(a, a, 12)
    assert e == (0 | (0x08000000 << 1) | (0x01000000 << 1))
    hp = bytes([1] * 14)                                    # homopolymer: all nine comparisons are 0 -> still (ZERO, ZERO)
    assert orc.lib().orc_umi_best9(hp, hp, 12) == e
    b = a[1:] + bytes([4])                                  # b shifted by one: window(a,+1) == window(b,0)
    e2 = orc.lib().orc_umi_best9(a, b, 12)
    assert e2 & 0xFFFFFF == 0 and (e2 >> 27) & 7 == 0b010 and (e2 >> 24) & 7 == 0b001    # pos1 ZERO, pos2 MINUSONE (first visited)


def test_umi_golden(orc):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "umi_*.npz"))):
        g = np.load(f)
        m, oo = orc.umi_matrix_batch(g["umis"], g["job_offsets"], int(g["umi_len"]))
        assert (m == g["matrix"]).all() and (oo == g["out_offsets"]).all()
