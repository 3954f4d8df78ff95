"""Formats either side of the hot path.  The read-name tests are pinned by the reference's own published examples
(/root/reference/README.md:400 and :452, quoted below verbatim)."""
import os

import numpy as np
import pytest

import __graft_entry__ as g

README_400 = ("_FWD_PS=566_PE=590_AE=619_bc=TCCGATCGTGCCAAGA_ed=0_ed_sec=2147483647_bcStart=618_bcEnd=603_rk=2987_"
              "X=AAAAAAAAAAAATGGCGTGTATTGTCTTGGCACGATCGGAAGA_Q=27.1")
README_452 = ("_REV_PS=1257_PE=1305_AE=1327_T=40_bc=GAGTGAGGTTGGGTAG_ed=1_ed_sec=2147483647_bcStart=1326_bcEnd=1311_rk=3883_"
              "X=AAAAAAAAAAACAAACCAAGTAACCAACCCAACCTCACTCAGA_Q=15.9")


def fmt():
    pkg = g.load_package()
    import importlib
    return importlib.import_module("sicelore_b200.formats")


def test_parse_readme_examples():
    F = fmt()
    with pytest.raises(ValueError):                      # as published (trimmed after Q=) the reference's own parser throws on the read-id field
        F.parse_read_name("b5c7a1f2-read" + README_400)
    a = F.parse_read_name("b5c7a1f2-read" + README_400 + "_")
    assert a == dict(reversed=False, polya_start=566, polya_end=590, adapter_end=619, ed=0, bc="TCCGATCGTGCCAAGA", ed_second=2147483647,
                     bc_start=618, bc_end=603, rank=2987, seq="AAAAAAAAAAAATGGCGTGTATTGTCTTGGCACGATCGGAAGA", mean_qv=float(np.float32(27.1)))
    b = F.parse_read_name("sp2" + README_452 + "_")
    assert b["reversed"] and b["tso_end"] == 40 and b["adapter_end"] == 1327 and b["bc"] == "GAGTGAGGTTGGGTAG" and b["ed"] == 1
    assert b["bc_start"] == 1326 and b["bc_end"] == 1311 and b["rank"] == 3883 and abs(b["mean_qv"] - 15.9) < 1e-6
    assert F.parse_read_name("read_without_tags") is None
    # the assignumis limit on the barcode ED drops the barcode block (FastqRecordExt.java:L450-L459)
    c = F.parse_read_name("x" + README_452 + "_1z", max_bc_ed=0)
    assert c["read_id"] == 71
    assert "bc" not in c and "rank" not in c and c["adapter_end"] == 1327


def test_write_reproduces_readme_examples():
    """rebuild the stranded read around the published X= string and write the extension back: identical up to the published end"""
    F = fmt()
    for ex, qv in ((README_400, 27.1), (README_452, 15.9)):
        p = F.parse_read_name("r" + ex + "_")
        ae, x = p["adapter_end"], p["seq"]
        begin = ae - 40 - 1                               # 3' geometry (L253): X = stranded[AE-41, AE+2)
        stranded = "C" * begin + x + "G" * 50
        # quality string whose mean formats as the published Q=; getMeanQV(quals, begin, end) skips begin - 1 chars and takes
        # end - begin + 1: the base before the X= range is averaged in as well (FastqRecordExt.java:L57-L59, L270)
        q_lo, q_hi, n = int(qv), int(qv) + 1, len(x) + 1
        n_hi = round((qv - q_lo) * n)
        quals = "I" * (begin - 1) + chr(33 + q_hi) * n_hi + chr(33 + q_lo) * (n - n_hi) + "I" * 50
        out = F.read_name_extension(p["reversed"], stranded, quals, adapter_end=ae, polya_start=p["polya_start"], polya_end=p["polya_end"],
                                    tso_end=p.get("tso_end"), bc=p["bc"], ed=p["ed"], ed_second=p["ed_second"], bc_start=p["bc_start"],
                                    bc_end=p["bc_end"], rank=p["rank"])
        assert out.startswith(ex + "_"), (out, ex)
        assert out.endswith(" cellBC=" + p["bc"])
        assert F.parse_read_name("r" + out.split(" ")[0]) == p                  # round trip


def test_barcode_geometry_of_readme_examples():
    """bc= is the reverse complement of X[-19:-3] up to `ed` edits; bcStart - bcEnd = 15 (3' reads count down towards the polyA)"""
    F = fmt()
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    a = F.parse_read_name("r" + README_400 + "_")
    assert rc(a["seq"][-19:-3]) == a["bc"] and a["bc_start"] - a["bc_end"] == 15 and a["bc_start"] == a["adapter_end"] - 1
    b = F.parse_read_name("r" + README_452 + "_")
    w = rc(b["seq"][-19:-3])
    assert sum(x != y for x, y in zip(w, b["bc"])) == b["ed"] == 1


def test_decimal_format():
    F = fmt()
    assert F._dec_format_1(27.1) == "27.1" and F._dec_format_1(27.0) == "27" and F._dec_format_1(9.96) == "10" and F._dec_format_1(0.25) == "0.2"


def test_assigned_tsv_round_trip(tmp_path):
    F = fmt()
    pkg = g.load_package()
    keys = np.array([pkg.pack_barcode(s) for s in ("CGGACTGTCTTGTACT", "AGCCTAAAGGGAAACA", "ACCCACTCAGTTCCCT", "TTTTTTTTTTTTTTTT")], dtype=np.uint64)
    counts = np.array([[118919, 21693, 5], [88046, 36699, 0], [8499, 112448, 0], [0, 0, 0]], dtype=np.int64)
    path = str(tmp_path / "BarcodesAssigned.tsv")
    assert F.write_assigned_tsv(path, keys, counts, 1) == 3                     # the unassigned barcode is not listed
    lines = open(path).read().split("\n")
    assert lines[0] == "Barcode\tn Reads with ED<=1 match\tED=0\tED=1"
    # the three rows the comment block of SelectValidCellBarcode.java:44-48 shows, in descending order of reads
    assert lines[1] == "CGGACTGTCTTGTACT\t140,612\t118,919\t21,693"
    assert lines[2] == "AGCCTAAAGGGAAACA\t124,745\t88,046\t36,699"
    assert lines[3] == "ACCCACTCAGTTCCCT\t120,947\t8,499\t112,448"
    rows = F.read_assigned_tsv(path)
    assert rows[0] == ("CGGACTGTCTTGTACT", 140612, 118919, 21693) and len(rows) == 3
    F.write_assigned_tsv(path, keys, counts, 2)
    assert open(path).readline() == "Barcode\tn Reads with ED<=2 match\tED=0\tED=1\tED=2\n"


@pytest.mark.parametrize("fname", ["ref_read_names.npz", "ref_read_names_wide.npz"])      # wide: oracle/make_ref_names_wide.py
def test_parser_matches_reference_bytecode(fname):
    """FastqRecordExt.getScanDatFromReadName run by the reference's own class files (oracle/minijvm.py, tests/golden/ref_read_names.npz): every
    parsed field, the absent / adapter-missing / read-id NumberFormatException outcomes, with and without the assignumis barcode-ED limit"""
    F = fmt()
    pkg = g.load_package()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname))
    outcomes = set()
    for name, lim, parsed in zip(z["name"], z["limit"], z["parsed"]):
        name, parsed = str(name), str(parsed)
        try:
            got = F.parse_read_name(name, max_bc_ed=None if lim < 0 else int(lim))
        except KeyError:
            assert parsed == "EXC:AdapterInfoNotFoundInReadException", (name, parsed)
            outcomes.add("adapter")
            continue
        except ValueError:
            assert parsed == "EXC:NumberFormatException", (name, parsed)
            outcomes.add("nfe")
            continue
        if got is None:
            assert parsed == "ABSENT"
            outcomes.add("absent")
            continue
        exp = dict(eval(parsed))
        want = {"Adapterresult.end": got.get("adapter_end"), "Forward.$ordinal": int(got["reversed"]), "read_id": got.get("read_id", 0)}
        if "tso_end" in got:
            want["TSOresult.end"] = got["tso_end"]
        if "polya_start" in got:
            want["PolyAResult.start"] = got["polya_start"]
        if "polya_end" in got:
            want["PolyAResult.end"] = got["polya_end"]
        for k, jk in (("ed", "editDistance"), ("ed_second", "editDistanceSecondBest"), ("bc_start", "start"), ("bc_end", "end"), ("rank", "rank")):
            if k in got:
                want["BarcodeResult." + jk] = got[k]
        if "bc" in got:
            want["BarcodeResult.barcodeseq"] = int(pkg.pack_barcode(got["bc"]))
        if "seq" in got:
            want["seq_len"] = len(got["seq"])
        if "mean_qv" in got:
            want["mean_qv"] = got["mean_qv"]
        assert want == exp, (name, want, exp)
        outcomes.add("ok" if "bc" in got else "ok_nobc")
    assert outcomes == {"adapter", "nfe", "absent", "ok", "ok_nobc"}


@pytest.mark.parametrize("fname", ["ref_written_names.npz", "ref_written_names_wide.npz"])
def test_writer_matches_reference_bytecode(fname):
    """FastqRecordExt.getRecordForWriting run by the reference's own class files (tests/golden/ref_written_names.npz): the whole read name for
    forward / reversed, 3' / 5' reads with and without polyA, TSO, barcode, second-best ED, rank, read id, and the adapter end too close to the
    read start for an X= slice"""
    F = fmt()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname))
    shapes = set()
    for name, stranded, quals, rev, five, rid, kw in zip(z["name"], z["stranded"], z["quals"], z["rev"], z["five"], z["read_id"], z["kw"]):
        kw = dict(eval(str(kw)))
        name = str(name)
        if name.startswith("EXC:"):                      # the reference itself throws (adapter end exactly at the start of the X= slice)
            assert name == "EXC:java/lang/IllegalArgumentException"
            with pytest.raises(ValueError):
                F.read_name_extension(bool(rev), str(stranded), str(quals), is5p=bool(five), read_id=None if rid < 0 else int(rid), **kw)
            shapes.add("throws")
            continue
        ext = F.read_name_extension(bool(rev), str(stranded), str(quals), is5p=bool(five), read_id=None if rid < 0 else int(rid), **kw)
        assert name == name.split("_")[0] + ext, (name, ext)
        shapes.add(("X=" in ext, "bc=" in ext, "T=" in ext, "PS=" in ext, rid >= 0))
        if ext == "":
            assert name == name.split("_")[0] and kw["adapter_end"] < 42          # "Beginrange inconsistent": the name stays bare
        if " cellBC=" in name:                           # the parser reads its own output back
            p = F.parse_read_name(name.split(" ")[0])
            assert p["bc"] == kw["bc"] and p["adapter_end"] == kw["adapter_end"] and p.get("read_id", 0) == max(int(rid), 0)
    assert len(shapes) >= 5 and ("throws" in shapes or "wide" not in fname)
