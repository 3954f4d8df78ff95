#!/usr/bin/env python
"""Real-jar validation kit, part 1: synthetic 3' reads as FASTQ + the -g barcode list ("the outstanding external validation").

    python baseline/make_fastq_kit.py <out_dir> [n_reads] [n_barcodes] [seed]

Writes  <out_dir>/fastq_pass/reads.fastq   reads in sequencing orientation, half of them reverse-complemented, layout of the stranded read
                                            (Jar/config.xml:111-113, README.md:376-452):  cDNA (300-600 nt) . A x 22-30 . revcomp(UMI 12) .
                                            revcomp(BC 16) . revcomp(adapter CTACACGACGCTCTTCCGATCT), Nanopore-like errors (sub 2 %, ins 1 %, del 2 %),
                                            10 % of the reads carry a random barcode, 1 % an N in the barcode
        <out_dir>/barcodes.tsv              the list for `scanfastq -g` (one barcode per line, "-1" suffix like Cellranger's barcodes.tsv)
        <out_dir>/truth.tsv                 read name, true barcode (or '-'), true UMI
Nothing here needs a GPU or the reference; run baseline/run_java_baseline.sh on a box with a JDK >= 13, then baseline/compare_with_jar.py."""
import os
import sys

import numpy as np

ADAPTER = "CTACACGACGCTCTTCCGATCT"
COMP = str.maketrans("ACGTN", "TGCAN")


def rc(s):
    return s.translate(COMP)[::-1]


def mutate(rng, s, p_sub=0.02, p_ins=0.01, p_del=0.02):
    out = []
    for ch in s:
        u = rng.random()
        if u < p_del:
            continue
        if u < p_del + p_sub:
            ch = "ACGT"[(("ACGT".index(ch) if ch in "ACGT" else 0) + int(rng.integers(1, 4))) % 4]
        out.append(ch)
        if rng.random() < p_ins:
            out.append("ACGT"[int(rng.integers(4))])
    return "".join(out)


def main():
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    n_bc = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(out, "fastq_pass"), exist_ok=True)
    rs = lambda k: "".join("ACGT"[i] for i in rng.integers(0, 4, k))
    barcodes = sorted({rs(16) for _ in range(n_bc)})
    with open(os.path.join(out, "barcodes.tsv"), "w") as f:
        for b in barcodes:
            f.write(b + "-1\n")
    with open(os.path.join(out, "fastq_pass", "reads.fastq"), "w") as fq, open(os.path.join(out, "truth.tsv"), "w") as tr:
        for i in range(n):
            true_bc = barcodes[int(rng.integers(len(barcodes)))] if rng.random() >= 0.10 else None
            bc = true_bc or rs(16)
            umi = rs(12)
            if rng.random() < 0.01:
                p = int(rng.integers(16))
                bc = bc[:p] + "N" + bc[p + 1:]
            stranded = rs(int(rng.integers(300, 600))) + "A" * int(rng.integers(22, 31)) + rc(umi) + rc(bc) + rc(ADAPTER) + rs(int(rng.integers(5, 30)))
            read = mutate(rng, stranded)
            if rng.random() < 0.5:
                read = rc(read)
            name = "read%07d" % i
            fq.write("@%s\n%s\n+\n%s\n" % (name, read, "".join(chr(33 + int(q)) for q in rng.integers(8, 35, len(read)))))
            tr.write("%s\t%s\t%s\n" % (name, true_bc or "-", umi))
    print("wrote %d reads and %d barcodes under %s" % (n, len(barcodes), out))


if __name__ == "__main__":
    main()
