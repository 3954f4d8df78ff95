"""Shared input builders for the parity tests (adversarial reads + whitelists that force hits at every ED level)."""
import random

import numpy as np

from oracle import pyref


def rcs(s):
    return s[::-1].translate(str.maketrans("ACGTNacgtn", "TGCANtgcan"))


def mutate(s, k, rng):
    s = list(s)
    for _ in range(k):
        op = rng.randrange(3)
        p = rng.randrange(len(s))
        if op == 0:
            s[p] = rng.choice("ACGT")
        elif op == 1:
            s.insert(p, rng.choice("ACGT"))
        else:
            del s[p]
    return "".join(s)


def adversarial(seed, three_prime, nreads, skew=False, nrand=300, dense=False, anchor=8):
    """32-char slices whose neighbourhoods are planted in the whitelist (0-3 edits, offsets -3..3), with N / IUPAC /
    lower-case / invalid characters sprinkled in; `skew` makes homopolymer-rich reads (duplicate mutants, visited-set
    logic); `dense` adds whole digit-group clusters (bucket overflow -> stash, many hits per bucket)."""
    rng = random.Random(seed)
    alpha = "AAAAACGT" if skew else "ACGT"
    reads, wl = [], set()
    for _ in range(nreads):
        r = "".join(rng.choice(alpha) for _ in range(32))
        if rng.random() < 0.15:
            p = rng.randrange(32)
            r = r[:p] + "N" + r[p + 1:]
        if rng.random() < 0.05:
            p = rng.randrange(32)
            r = r[:p] + rng.choice("RYKM-X*acgtn") + r[p + 1:]
        reads.append(r)
        for _ in range(rng.randrange(0, 8)):
            o = rng.randrange(-3, 4)
            seg = rcs(r[max(0, anchor + o - 6):anchor + o + 16]) if three_prime else r[anchor + o:anchor + o + 22]
            seg = "".join(c if c in "ACGT" else "A" for c in seg.upper())
            m = mutate(seg, rng.randrange(0, 4), rng)[:16]
            if len(m) == 16:
                wl.add(pyref.pack(m))
    if dense:
        for r in reads[:20]:
            o = rng.randrange(-2, 3)
            seg = rcs(r[anchor + o:anchor + o + 16]) if three_prime else r[anchor + o:anchor + o + 16]
            seg = "".join(c if c in "ACGT" else "A" for c in seg.upper())
            k = pyref.pack(seg)
            g = rng.randrange(4)
            sh = 24 - 8 * g
            for pat in range(256):
                if rng.random() < 0.6:
                    wl.add((k & ~(0xFF << sh)) | (pat << sh))
            k2 = ((k << 2) | rng.randrange(4)) & 0xFFFFFFFF
            for pat in range(256):
                if rng.random() < 0.3:
                    wl.add((k2 & ~(0xFF << sh)) | (pat << sh))
    for _ in range(nrand):
        wl.add(rng.getrandbits(32))
    wl = sorted(wl)
    rng.shuffle(wl)
    slices = np.array([np.frombuffer(r.encode(), dtype=np.uint8) for r in reads])
    anchors = np.full(len(reads), anchor, dtype=np.int32)
    return reads, slices, anchors, np.array(wl, dtype=np.uint64)


def umi_jobs(seed, umi_len=12, n_jobs=60, max_n=40):
    rng = np.random.default_rng(seed)
    codes = np.array([1, 2, 4, 8], dtype=np.uint8)
    sizes = rng.integers(1, max_n, size=n_jobs)
    rows = []
    for n in sizes:
        base = codes[rng.integers(0, 4, size=(3, umi_len + 2))]
        for _ in range(n):
            u = base[rng.integers(0, 3)].copy()
            for _ in range(rng.integers(0, 4)):
                op = rng.integers(0, 3)
                p = rng.integers(0, umi_len + 2)
                if op == 0:
                    u[p] = codes[rng.integers(0, 4)]
                elif op == 1:
                    u = np.concatenate([u[:p], [codes[rng.integers(0, 4)]], u[p:]])[:umi_len + 2]
                else:
                    u = np.concatenate([u[:p], u[p + 1:], [codes[rng.integers(0, 4)]]])
            if rng.random() < 0.1:
                u[rng.integers(0, umi_len + 2)] = 15
            rows.append(u)
    umis = np.zeros((len(rows), 16), dtype=np.uint8)
    umis[:, :umi_len + 2] = np.array(rows)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return umis, offs


def used_list(seed, n_cells, skew=False, n_children=6):
    """A pass-1 style used-barcode list: `n_cells` random barcodes plus 0..n_children-1 variants of each at 1-3 edits
    (sequencing-error children that the collision tester is there to find); `skew` = homopolymer-rich (duplicate mutants)."""
    rng = random.Random(seed)
    alpha = "AAAAACGT" if skew else "ACGT"
    wl = set()
    for _ in range(n_cells):
        s = "".join(rng.choice(alpha) for _ in range(16))
        wl.add(pyref.pack(s))
        for _ in range(rng.randrange(0, n_children)):
            m = (mutate(s, rng.randrange(1, 4), rng) + "".join(rng.choice("ACGT") for _ in range(3)))[:16]
            wl.add(pyref.pack(m))
    wl = sorted(wl)
    rng.shuffle(wl)
    return np.array(wl, dtype=np.uint64)
