#!/usr/bin/env python
"""Real-jar validation kit, part 3: diff the reference's own output against this repository's barcode path.

    python baseline/compare_with_jar.py --scan-dir <scanfastq output dir> --list <barcodes.tsv> --ed 2 [--oracle] [--limit N]

Every record of <scan-dir>/passed/* (stranded reads) and <scan-dir>/failed/* (reads in sequencing orientation; stranded here from the _FWD / _REV
mark) whose name carries the adapter end (AE=) is replayed: the 32-byte slice around the adapter end and the anchor are built exactly as the
JNI shim builds them, pushed through slr_bc_assign (default) or the CPU oracle (--oracle, for a JDK box without a GPU), and
bc= / ed= / ed_sec= / bcStart= / bcEnd= of the jar's read name are compared with the record.  Prints a parity report; exit code 1 on any mismatch."""
import argparse
import glob
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def fastq_records(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        while True:
            name = f.readline()
            if not name:
                return
            seq = f.readline().strip()
            f.readline()
            f.readline()
            yield name[1:].split()[0], seq                        # the read name ends at the first blank (' cellBC=' is the FASTQ comment)


def build_slice(stranded, adapterpos, three_prime=True):
    """slice start / anchor of INTEGRATION.md 2: 0-based start of the offset-0 window, 8 bytes of left flank, at most 32 bytes"""
    ws0 = adapterpos - 16 - 1 if three_prime else adapterpos
    s0 = max(0, ws0 - 8)
    piece = stranded[s0:s0 + 32].encode()
    sl = np.zeros(32, dtype=np.uint8)
    sl[:len(piece)] = np.frombuffer(piece, dtype=np.uint8)
    return sl, ws0 - s0, len(piece)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan-dir", required=True)
    ap.add_argument("--list", required=True)
    ap.add_argument("--ed", type=int, default=2)
    ap.add_argument("--plusminus", type=int, default=2)
    ap.add_argument("--five-prime", action="store_true")
    ap.add_argument("--oracle", action="store_true", help="CPU oracle instead of the GPU library")
    ap.add_argument("--limit", type=int, default=0)
    a = ap.parse_args()
    import __graft_entry__ as g
    pkg = g.load_package()
    import importlib
    fmt = importlib.import_module("sicelore_b200.formats")
    keys = pkg.read_whitelist(a.list)
    tp = not a.five_prime
    names, slices, anchors, lens, want = [], [], [], [], []
    for sub, stranded_already in (("passed", True), ("failed", False)):
        for path in sorted(glob.glob(os.path.join(a.scan_dir, sub, "*"))):
            for name, seq in fastq_records(path):
                d = fmt.parse_read_name(name)
                if d.get("adapter_end") is None:
                    continue
                stranded = seq if stranded_already or not d.get("reversed") else seq.encode().translate(COMP)[::-1].decode()
                sl, anc, ln = build_slice(stranded, int(d["adapter_end"]), tp)
                names.append(name); slices.append(sl); anchors.append(anc); lens.append(ln)
                want.append((d.get("bc"), d.get("ed"), d.get("ed_second"), d.get("bc_start"), d.get("bc_end"), int(d["adapter_end"])))
                if a.limit and len(names) >= a.limit:
                    break
    if not names:
        print("no read with an adapter end found under", a.scan_dir)
        return 2
    slices = np.stack(slices); anchors = np.array(anchors, dtype=np.int32); lens = np.array(lens, dtype=np.int32)
    rank = np.arange(1, len(keys) + 1, dtype=np.int32)
    if a.oracle:
        from oracle import orc
        res, _ = orc.assign_barcode_batch(orc.BarcodeSet(keys, rank), slices, anchors, a.ed, a.plusminus, tp)   # (slices are zero-padded: short reads throw alike)
    else:
        ctx = pkg.Context(0)
        res = pkg.Parser(ctx, pkg.BarcodesMapForBCfinding(ctx, keys, rank), a.ed, a.plusminus, tp).assign_barcodes(slices, anchors, lens=lens)
    bad = n_assigned = 0
    for i, (bc, ed, ed2, bs, be, apos) in enumerate(want):
        r = res[i]
        if bc is None:
            ok = not (r["flags"] & 1)
        else:
            st, en = pkg.Parser.barcode_positions(res[i:i + 1], apos, tp)
            ok = bool(r["flags"] & 1) and fmt.unpack2bit(int(r["bc"])) == bc and int(r["ed"]) == int(ed) and int(r["ed_second"]) == int(ed2) \
                and int(st[0]) == int(bs) and int(en[0]) == int(be)
            n_assigned += 1
        if not ok:
            bad += 1
            if bad <= 10:
                print("MISMATCH", names[i][:60], "jar:", want[i], "here:", dict(bc=fmt.unpack2bit(int(r["bc"])), ed=int(r["ed"]), ed_sec=int(r["ed_second"]),
                                                                                flags=int(r["flags"])))
    print("compared %d reads with an adapter end (%d assigned by the jar): %d mismatches  [%s, list of %d, --bcEditDistance %d]"
          % (len(want), n_assigned, bad, "CPU oracle" if a.oracle else "libsicelore_gpu", len(keys), a.ed))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
