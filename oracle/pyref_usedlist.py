"""TEST INFRASTRUCTURE — Python restatement of the hand-over between scanfastq's two passes: UsedBarcodesListData.finalizeData
(F!com/rw/nanoporereadscanner/analyzers/UsedCellBCListGenerator$UsedBarcodesListData.class, UsedCellBCListGenerator.java:L359-L363, L379-L406) and
BarcodeDatasetColissionTester.generateColissionMergedBCmap (F!…/BarcodeDatasetColissionTester.class, …java:L158-L203), which consume the
pass-1 counts and the collision tester's Matches (slr_bc_collide) and produce the used-barcode list of pass 2.  The product's implementation is
csrc/slr_usedlist.cpp; pinned by oracle/make_ref_usedlist.py -> tests/golden/ref_usedlist.npz."""
import numpy as np

F_ORDER_UNPIN = 1      # a java.util.HashMap bin reached 9 entries: the JDK turns it into a tree (or resizes early), iteration order not reproduced
F_RANK_TIES = 2        # equal counts among the kept barcodes: their ranks follow fastutil's table order in the reference


def filter_low_counts(counts, record_count):
    """filterLowCounts with the cutoff of finalizeData: count > 2.0f * recordCount / 5000000.0f (float arithmetic) and count > 1"""
    cutoff = np.float32(np.float32(2.0) * np.float32(int(record_count))) / np.float32(5000000.0)
    c = np.asarray(counts, dtype=np.int64)
    return (c.astype(np.float32) > cutoff) & (c > 1)


def long_hash_bucket(v, cap):
    """java.util.HashMap bucket of a Long key: spread(Long.hashCode(v)) & (cap - 1)"""
    v &= 0xFFFFFFFFFFFFFFFF
    h = (v ^ (v >> 32)) & 0xFFFFFFFF
    return (h ^ (h >> 16)) & (cap - 1)


def hashmap_key_order(keys):
    """iteration order of a java.util.HashMap<Long, ?> filled with put() in the given order (distinct keys): bins in index order, a bin in
    insertion order (a resize splits a bin without reordering it).  Growth: doubled when size exceeds 0.75 x capacity, and — below 64 —
    when a bin receives its 9th entry (treeifyBin resizes instead).  Second value: a bin reached 9 entries at capacity >= 64 (tree bin)."""
    cap, size, unpinned = 16, 0, False
    cnt = {}
    done = []

    def recount():
        cnt.clear()
        for k in done:
            b = long_hash_bucket(k, cap)
            cnt[b] = cnt.get(b, 0) + 1
    for k in keys:
        b = long_hash_bucket(k, cap)
        done.append(k)
        cnt[b] = cnt.get(b, 0) + 1
        if cnt[b] >= 9:
            if cap < 64:
                cap *= 2
                recount()
            else:
                unpinned = True
        size += 1
        if size > 0.75 * cap:
            cap *= 2
            recount()
    order = sorted(range(len(done)), key=lambda i: (long_hash_bucket(done[i], cap), i))
    return [done[i] for i in order], unpinned


def merge_collisions(barcodes, counts, collide, min_count_fold, merge_ed, cells_fold, visiting="hashmap", lazy=True, strict=True):
    """generateColissionMergedBCmap + the ranks of WorkerReadscanner.java:L264-L269.  collide: COLLIDE_RESULT per barcode (slr_bc_collide of the
    list against itself).  Returns (keep mask, rank (1-based, 0 = dropped), flags).  The keyword arguments switch single behaviours of the
    reference OFF (tests/test_usedlist.py shows that the vectors notice): visiting="insertion" walks the map in count order instead of the JDK's
    bin order, lazy=False lets removed barcodes remove others, strict=False removes colliders with count == cutoff too."""
    barcodes = [int(b) for b in barcodes]
    counts = [int(c) for c in counts]
    n = len(barcodes)
    index = {b: i for i, b in enumerate(barcodes)}
    assert len(index) == n
    # L164-L183: barcodes with a non-empty Matches, by count descending (stable), each with the set of its much smaller colliders
    with_matches = [i for i in range(n) if collide[i]["valid"]]
    with_matches.sort(key=lambda i: -counts[i])
    to_merge = {}
    for i in with_matches:
        cutoff = counts[i] // min_count_fold
        to_merge[barcodes[i]] = [index[int(collide[i]["bc"][e])] for e in range(2)
                                 if collide[i]["valid"] >> e & 1 and e + 1 <= merge_ed and
                                 (counts[index[int(collide[i]["bc"][e])]] < cutoff or (not strict and counts[index[int(collide[i]["bc"][e])]] == cutoff))]
    order, unpinned = hashmap_key_order(list(to_merge.keys()))
    if visiting == "insertion":
        order = list(to_merge.keys())
    alive = [True] * n
    for b in order:                                            # L186-L195: HashMap order; a barcode that was itself removed removes nobody
        if alive[index[b]] or not lazy:
            for c in to_merge[b]:
                alive[c] = False
    live = [counts[i] for i in range(n) if alive[i]]
    if not live:
        raise ValueError("java.util.NoSuchElementException (BarcodeDatasetColissionTester.java:L197: empty list)")
    min_counts = max(live) // cells_fold                       # L197-L198
    keep = np.array([alive[i] and counts[i] >= min_counts for i in range(n)], dtype=bool)
    kept = sorted(np.nonzero(keep)[0].tolist(), key=lambda i: -counts[i])
    rank = np.zeros(n, dtype=np.int32)
    for r, i in enumerate(kept):
        rank[i] = r + 1
    ties = len({counts[i] for i in kept}) != len(kept)
    return keep, rank, (F_ORDER_UNPIN if unpinned else 0) | (F_RANK_TIES if ties else 0)
