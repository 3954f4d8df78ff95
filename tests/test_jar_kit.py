"""The real-jar validation kit (baseline/): FASTQ + list writer, and the comparator from the jar's read names to slr_bc_result.
No JVM here, so the "jar output" of these tests is synthesised: CPU oracle for the assignment + formats.read_name_extension for the names
(both pinned to the reference's bytecode elsewhere: test_ref_vectors.py, test_formats.py).  What is under test is the kit's own plumbing —
slice / anchor reconstruction from AE=, strand handling of failed/ reads, position arithmetic, mismatch detection."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import __graft_entry__ as g
from oracle import orc

pkg = g.load_package()
import importlib
fmt = importlib.import_module("sicelore_b200.formats")
import compare_with_jar as cmpjar

COMP = str.maketrans("ACGTN", "TGCAN")


def _fake_scan(tmp, kit, ed):
    """what scanfastq would leave behind, for reads whose adapter we place ourselves: passed/ (stranded, assigned) and failed/ (as sequenced)"""
    keys = pkg.read_whitelist(os.path.join(kit, "barcodes.tsv"))
    bset = orc.BarcodeSet(keys, np.arange(1, len(keys) + 1, dtype=np.int32))
    os.makedirs(os.path.join(tmp, "passed")); os.makedirs(os.path.join(tmp, "failed"))
    recs = list(cmpjar.fastq_records(os.path.join(kit, "fastq_pass", "reads.fastq")))
    adapter_rc = "AGATCGGAAGAGCGTCGTGTAG"
    n_pass = n_fail = 0
    with open(os.path.join(tmp, "passed", "p.fastq"), "w") as fp, open(os.path.join(tmp, "failed", "f.fastq"), "w") as ff:
        for name, seq in recs:
            rev = False
            k = seq.find(adapter_rc[:10])
            if k < 0:
                seq2 = seq.translate(COMP)[::-1]
                k = seq2.find(adapter_rc[:10])
                if k < 0:
                    continue
                rev, stranded = True, seq2
            else:
                stranded = seq
            ae = k                                                              # 1-based position of the last barcode base + 1 = adapter start; the
            sl, anc, ln = cmpjar.build_slice(stranded, ae + 1, True)             # scanner reports the adapter END on the stranded 3' read = k + 1 - 1 + 1
            res, _ = orc.assign_barcode_batch(bset, sl[None, :], np.array([anc], dtype=np.int32), ed, 2, True)
            r = res[0]
            q = "I" * len(seq)
            if r["flags"] & 1:
                st, en = pkg.Parser.barcode_positions(res[:1], ae + 1, True)
                ext = fmt.read_name_extension(rev, stranded, q, adapter_end=ae + 1, polya_start=max(1, ae - 60), polya_end=max(2, ae - 30), bc=int(r["bc"]),
                                              ed=int(r["ed"]), ed_second=int(r["ed_second"]), bc_start=int(st[0]), bc_end=int(en[0]), rank=int(r["rank"]),
                                              read_id=n_pass)
                fp.write("@%s%s\n%s\n+\n%s\n" % (name, ext, stranded, q)); n_pass += 1
            else:
                ext = fmt.read_name_extension(rev, stranded, q, adapter_end=ae + 1, read_id=n_fail)
                ff.write("@%s%s\n%s\n+\n%s\n" % (name, ext, seq, q)); n_fail += 1
    return n_pass, n_fail


@pytest.fixture(scope="module")
def kit(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("kit"))
    subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "make_fastq_kit.py"), d, "600", "300", "5"], check=True, stdout=subprocess.DEVNULL)
    return d


def test_kit_writer(kit):
    assert sum(1 for _ in open(os.path.join(kit, "barcodes.tsv"))) == len(pkg.read_whitelist(os.path.join(kit, "barcodes.tsv"))) > 250
    recs = list(cmpjar.fastq_records(os.path.join(kit, "fastq_pass", "reads.fastq")))
    assert len(recs) == 600 and all(len(s) > 300 for _, s in recs)
    assert sum(1 for _ in open(os.path.join(kit, "truth.tsv"))) == 600


def _run(scan, kit, *extra):
    return subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "compare_with_jar.py"), "--scan-dir", scan, "--list",
                           os.path.join(kit, "barcodes.tsv"), "--ed", "2", *extra], capture_output=True, text=True)


def test_comparator_oracle_self_test(kit, tmp_path):
    scan = str(tmp_path / "scan")
    os.makedirs(scan)
    n_pass, n_fail = _fake_scan(scan, kit, 2)
    assert n_pass > 200 and n_fail > 10
    r = _run(scan, kit, "--oracle")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compared %d reads" % (n_pass + n_fail) in r.stdout and " 0 mismatches" in r.stdout
    # a tampered name must be reported: flip ed= of the first passed read
    p = os.path.join(scan, "passed", "p.fastq")
    lines = open(p).read().split("\n")
    import re
    lines[0] = re.sub(r"_ed=(\d)_", lambda m: "_ed=%d_" % (int(m.group(1)) + 1), lines[0], count=1)
    open(p, "w").write("\n".join(lines))
    r = _run(scan, kit, "--oracle")
    assert r.returncode == 1 and " 1 mismatches" in r.stdout


@pytest.mark.gpu
def test_comparator_gpu(kit, tmp_path):
    scan = str(tmp_path / "scan")
    os.makedirs(scan)
    n_pass, n_fail = _fake_scan(scan, kit, 2)
    r = _run(scan, kit)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout and "libsicelore_gpu" in r.stdout, r.stdout + r.stderr
