"""Wide bytecode pin of the dominant kernel: Parser.assignBarcode (Parser.java:L195-L315) run BY THE REFERENCE'S OWN CLASS FILES
(oracle/minijvm.py) on reads of the bench generator at --bcEditDistance 2, frozen in tests/golden/ref_assign_wide.npz.

    python oracle/make_ref_assign_wide.py [n_bench_reads] [n_constructed] [processes]

Part 1 — bench reads.  The first n reads of bench.py's bc3m_ed2 workload (synth_reads(whitelist 3 000 000 / seed 3 000 000, seed 2)).
The interpreter cannot hold a 3 M-entry map per process cheaply, and does not have to: the reference only asks its search set
`contains(mutant)`, and every mutant it can ever probe lies within two engine operations of one of the read's five windows.  For every
read the script enumerates a SUPERSET of that neighbourhood (all substitutions, insertions and deletions with every appended base, twice,
without the engine's skips) and intersects it with the 3 M list; the union U of those hits over all reads (plus a few hundred random list
members) is the search set of the run.  U and the full list agree on every key the reference can probe for these reads, so the frozen
results are the reference's results ON THE FULL 3 M LIST — the GPU test checks both tables.  Ranks are the list positions + 1 (as bench.py).
Part 2 — constructed reads with many merged OneMatch entries: periodic reads (identical windows at several offsets -> one HashMap bin),
planted list members at ED 0 / 1 / 2 of several windows: the 16 -> 32 resize above 12 entries and the 9-node chain resize of the merged
java.util.HashSet (minijvm.JdkHashSet), +-1 ... +-4 windows.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_ref_vectors as V  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_assign_wide.npz")
L = 16
M32 = 0xFFFFFFFF


def neighbours1(w):
    """all single engine operations on 32-bit windows w (uint64 array, low 32 bits): SUB 4 x 16, INS 4 x 15, DEL (every appended base) 4 x 15"""
    out = []
    w = w.astype(np.uint64)
    for p in range(16):
        sh = np.uint64(2 * (15 - p))
        below = (np.uint64(1) << sh) - np.uint64(1)
        for b in range(4):
            out.append((w & ~(np.uint64(3) << sh) & np.uint64(M32)) | (np.uint64(b) << sh))
            if p < 15:
                out.append(((w & ~below) | ((w & below) >> np.uint64(2)) | ((np.uint64(b) << sh) >> np.uint64(2))) & np.uint64(M32))
                below2 = (below << np.uint64(2)) | np.uint64(3)
                out.append(((w & ~below2) | ((w << np.uint64(2)) & below2) | np.uint64(b)) & np.uint64(M32))
    return np.unique(np.concatenate(out))


def windows_of(slice_bytes, anchor, pm, three_prime=True):
    comp = {65: 3, 71: 2, 67: 1, 84: 0}          # complement in the reference's 2-bit code A=0 G=1 C=2 T=3
    code = {65: 0, 71: 1, 67: 2, 84: 3}
    ws = []
    for o in range(-pm, pm + 1):
        s = anchor + o
        if s < 0 or s + 16 > len(slice_bytes):
            continue
        chars = slice_bytes[s:s + 16]
        if any(c not in code for c in chars):
            continue
        v = 0
        if three_prime:
            for c in reversed(chars):
                v = (v << 2) | comp[c]
        else:
            for c in chars:
                v = (v << 2) | code[c]
        ws.append(v)
    return ws


_vm = None
_P = {}


def _worker(job):
    """one read through the reference's assignBarcode; the search set / parameters are shared per (list id, ed, pm, three_prime)"""
    global _vm
    kind, read, ap, ed, pm, tp, keys, ranks = job
    if _vm is None:
        _vm = V.J.VM(V.JARS)
    pk = (kind, ed, pm, tp) if kind == "U" else None
    if pk is not None and pk in _P:
        P = _P[pk]
    else:
        P = V.make_parser(_vm, keys, ranks, ed, pm, tp)
        if pk is not None:
            _P[pk] = P
    P.f["assignedBarcodes2ndPass"].v.clear()
    t0 = time.time()
    res = V.run_assign(_vm, P, read, ap)
    return res, time.time() - t0


def main():
    n_bench = int(sys.argv[1]) if len(sys.argv) > 1 else 520
    n_cons = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    import __graft_entry__ as g
    pkg = g.load_package()
    t0 = time.time()
    wl = pkg.synth_whitelist(3_000_000, 3_000_000)
    slices, anchor, _ = pkg.synth_reads(wl, n_bench, seed=2)
    wl_sorted = np.sort(wl)
    order = np.argsort(wl, kind="stable")
    rng = np.random.default_rng(4711)
    hit = [wl[rng.integers(0, len(wl), 300)]]
    for i in range(n_bench):
        ws = windows_of(bytes(slices[i]), int(anchor[i]), 2)
        if not ws:
            continue
        n1 = neighbours1(np.array(ws, dtype=np.uint64))
        n2 = neighbours1(n1)
        cand = np.unique(np.concatenate([np.array(ws, dtype=np.uint64), n1, n2]))
        pos = np.searchsorted(wl_sorted, cand)
        ok = wl_sorted[np.minimum(pos, len(wl_sorted) - 1)] == cand
        hit.append(cand[ok])
    U = np.unique(np.concatenate(hit))
    pos = np.searchsorted(wl_sorted, U)
    U_rank = (order[pos] + 1).astype(np.int32)
    print("bench part: %d reads, search set U = %d of the 3 M list members (%.1f s)" % (n_bench, len(U), time.time() - t0), flush=True)
    keys_U, ranks_U = [int(k) for k in U], [int(r) for r in U_rank]
    jobs = [("U", bytes(slices[i]).decode("latin-1"), int(anchor[i]) + 17, 2, 2, True, keys_U, ranks_U) for i in range(n_bench)]

    # ---- part 2: constructed reads ------------------------------------------------------------------------------------------
    comp = str.maketrans("ACGT", "TGCA")
    cons = []
    for t in range(n_cons):
        period = [1, 2, 2, 4, 3, 16][t % 6]
        unit = V.rseq(rng, period)
        if period == 1 and t % 12 == 0:
            unit = "A"
        read = (unit * 64)[:56]
        if t % 5 == 4:                                           # break the period in the middle of the window span
            p = int(rng.integers(20, 36))
            read = read[:p] + V.rseq(rng, 1) + read[p + 1:]
        pm = [2, 2, 3, 4, 1][t % 5]
        tp = t % 4 != 3
        ap = int(rng.integers(26, 32)) if tp else int(rng.integers(8, 14))
        keys = {V.pack(V.rseq(rng, L)) for _ in range(20)}
        for o in range(-pm, pm + 1):                             # plant members at ED 0 / 1 / 2 of every window
            if tp:
                b, e = ap - L - 1 + o, ap - 1 + o
                w = read[b:e][::-1].translate(comp) if 0 <= b and e <= len(read) else None
            else:
                b = ap + o
                w = read[b:b + L] if b + L <= len(read) else None
            if w is None or len(w) != L:
                continue
            for d in (0, 1, 2):
                if rng.random() < (0.9 if t % 3 else 0.6):
                    keys.add(V.pack((V.mutate(rng, w, d) + V.rseq(rng, 4))[:L]))
        keys = sorted(keys)
        cons.append(dict(read=read, ap=ap, ed=2, pm=pm, tp=tp, keys=keys))
        jobs.append(("C%d" % t, read, ap, 2, pm, tp, keys, list(range(1, len(keys) + 1))))

    with Pool(procs) as pool:
        out = []
        for k, r in enumerate(pool.imap(_worker, jobs, chunksize=1)):
            out.append(r)
            if k % 20 == 0:
                print("  %d / %d reads, %.0f s, last read %.1f s" % (k, len(jobs), time.time() - t0, r[1]), flush=True)
    res = [r[0] for r in out]

    def pack_rows(rs):
        status = np.array([2 if isinstance(r, str) else (0 if r is None else 1) for r in rs], dtype=np.int32)
        rows = np.array([list(r) if isinstance(r, tuple) else [0] * 7 for r in rs], dtype=np.uint64).reshape(-1, 7)
        return status, rows
    sb, rb = pack_rows(res[:n_bench])
    sc, rc = pack_rows(res[n_bench:])
    koff = np.cumsum([0] + [len(c["keys"]) for c in cons]).astype(np.int64)
    np.savez_compressed(OUT, slices=slices, anchor=anchor, U=U, U_rank=U_rank, status=sb, result=rb,
                        c_read=np.array([c["read"] for c in cons]), c_ap=np.array([c["ap"] for c in cons], dtype=np.int32),
                        c_pm=np.array([c["pm"] for c in cons], dtype=np.int32), c_tp=np.array([c["tp"] for c in cons], dtype=np.int32),
                        c_keys=np.array([k for c in cons for k in c["keys"]], dtype=np.uint64), c_key_offsets=koff, c_status=sc, c_result=rc)
    print("bench reads: unassigned/assigned/exception", np.bincount(sb, minlength=3), " constructed:", np.bincount(sc, minlength=3),
          " %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
