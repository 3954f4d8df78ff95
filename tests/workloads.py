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


# ---- Illumina-guided search ---------------------------------------------------------------------------
G_BASES = "AGCT"                       # 2-bit order of the reference: A=0 G=1 C=2 T=3


def g_pack(s):
    v = 0
    for ch in s:
        v = (v << 2) | G_BASES.index(ch)
    return v


def g_unpack(v, L):
    return "".join(G_BASES[(int(v) >> (2 * (L - 1 - i))) & 3] for i in range(L))


def guided(seed, L, nq, max_err, pm, post_len, bc_flavour, group_sizes=(0, 1, 2, 5, 12, 40), skew=False, n_all=30, n_empty=60,
           special=True):
    """Synthetic guided-search batch: candidate groups (CSR), optionally the two global lists of the BC flavour, and 32-byte
    stranded slices whose windows sit 0..max_err edits from a candidate, shifted by up to pm+1 bases, with N / IUPAC /
    lower-case / invalid characters sprinkled in and homopolymer-rich windows (`skew`: duplicate mutants -> visited set).
    Returns dict(group_keys, group_offsets, all_keys, empty_keys, slices, anchor, group_id, slice_len)."""
    rng = random.Random(seed)
    alpha = "AAAAAGCT" if skew else "AGCT"
    rseq = lambda n, a=alpha: "".join(rng.choice(a) for _ in range(n))
    groups = []
    for m in group_sizes:
        g = {g_pack(rseq(L)) for _ in range(m)}
        if special and m >= 5:
            g.add(g_pack("T" * L))                      # all-T: collides with the empty-slot marker of the device tables when L = 16
            g.add(g_pack("A" * L))
        groups.append(sorted(g))
    allk = emptyk = None
    if bc_flavour:
        allk = sorted({k for g in groups for k in g} | {g_pack(rseq(L)) for _ in range(n_all)})
        emptyk = sorted({g_pack(rseq(L)) for _ in range(n_empty)} | ({g_pack("T" * L)} if special else set()))
    slice_len = min(32, 2 * pm + L + post_len + 2)
    slices = np.zeros((nq, 32), dtype=np.uint8)
    anchor = np.zeros(nq, dtype=np.int32)
    gid = np.zeros(nq, dtype=np.int32)
    for q in range(nq):
        g = rng.randrange(len(groups))
        gid[q] = g if rng.random() > 0.03 else rng.choice([-1, len(groups) + 3])
        r = rng.random()
        if groups[g] and r < 0.7:
            true = g_unpack(rng.choice(groups[g]), L)
        elif bc_flavour and r < 0.85:
            true = g_unpack(rng.choice(allk), L)
        elif bc_flavour and r < 0.95:
            true = g_unpack(rng.choice(emptyk), L)
        else:
            true = rseq(L)
        mid = mutate(true, rng.randrange(0, max_err + 2), rng).replace("A", rng.choice("Aa"), 1)
        a = pm + (1 if 2 * pm + L + post_len + 1 <= 32 else 0)
        lead = a + rng.choice([-1, 0, 0, 0, 1]) if pm > 0 else a
        s = (rseq(max(lead, 0), "AGCT") + mid + rseq(32, "AGCT"))[:32]
        x = rng.random()
        if x < 0.04:
            p = rng.randrange(32)
            s = s[:p] + "N" + s[p + 1:]
        elif x < 0.07:
            p = rng.randrange(32)
            s = s[:p] + rng.choice("RYKMSWBDHV") + s[p + 1:]
        elif x < 0.08:
            p = rng.randrange(32)
            s = s[:p] + rng.choice("X*@") + s[p + 1:]
        slices[q] = np.frombuffer(s.encode(), dtype=np.uint8)
        anchor[q] = a if rng.random() > 0.02 else rng.choice([0, 1, 32 - L])      # windows / post leaving the slice -> exception
    return dict(group_keys=np.array([k for g in groups for k in g], dtype=np.uint64),
                group_offsets=np.cumsum([0] + [len(g) for g in groups]).astype(np.int64),
                all_keys=None if allk is None else np.array(allk, dtype=np.uint64),
                empty_keys=None if emptyk is None else np.array(emptyk, dtype=np.uint64),
                slices=slices, anchor=anchor, group_id=gid, slice_len=slice_len, groups=groups)
