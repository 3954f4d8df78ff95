"""Clustering of the small jobs + UMI assignment: ClusterOneHierarchical.call
(F!com/rw/umifinder/analyzers/clustering/ClusterOneHierarchical.class, ClusterOneHierarchical.java:L61-L217) with LingPipe's
CompleteLinkClusterer (A!com/aliasi/cluster/CompleteLinkClusterer.class, CompleteLinkClusterer.java:L146-L237), OneUmiCluster.setClusterCenter
and the per-read values of ClusterOneBase.setSamflagsAndStatsForClustered.

CPU: the C oracle against the object-for-object Python restatement, and against the expectations of LingPipe's OWN unit tests, which ship
inside Jar/lib/Aliasi_ClusteringLib-1.0.jar (com/aliasi/test/unit/cluster/CompleteLinkClustererTest.testOne, SingleLinkClustererTest).
GPU: slr_umi_assign / slr_umi_assign_dev / slr_umi_session_assign through the C ABI against the oracle (bit-exact records)."""
import numpy as np
import pytest

from oracle import pyref


def random_packed(rng, n, mode):
    """a job's packed matrix: ED 0..5 in the low byte, one-hot shift flags, lower triangle = transposed copy (getTransposedCopy)"""
    if mode == 0:                                            # blocks of near-identical reads (what UMI data looks like) + noise
        lab = rng.integers(0, max(1, n // 3), n)
        base = np.where(lab[:, None] == lab[None, :], rng.integers(0, 3, (n, n)), rng.integers(2, 6, (n, n)))
    elif mode == 1:                                          # exact cliques: many equal costs, every merge order gives the same partition
        lab = rng.integers(0, max(1, n // 4), n)
        base = np.where(lab[:, None] == lab[None, :], rng.integers(0, 2, (n, n)), 5)
    else:
        pc = float(rng.choice([0.05, 0.2, 0.5, 0.9]))
        base = np.where(rng.random((n, n)) < pc, rng.integers(0, 3, (n, n)), rng.integers(3, 6, (n, n)))
    e = np.triu(base, 1)
    e = e + e.T
    p1, p2 = rng.integers(0, 3, (n, n)), rng.integers(0, 3, (n, n))
    up = e | (0x08000000 << p1) | (0x01000000 << p2)
    lo = e | (0x08000000 << p2.T) | (0x01000000 << p1.T)
    packed = np.where(np.arange(n)[:, None] <= np.arange(n)[None, :], up, lo)
    np.fill_diagonal(packed, 0x10000000 | 0x02000000)
    return packed.astype(np.int32)


def make_batch(rng, n_jobs, max_n=40, big_every=10):
    mats, sizes = [], []
    for t in range(n_jobs):
        n = int(rng.integers(1, max_n)) if t % big_every else int(rng.integers(33, 101))
        mats.append(random_packed(rng, n, t % 3).ravel())
        sizes.append(n)
    offs = np.zeros(n_jobs + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    oo = np.zeros(n_jobs + 1, dtype=np.int64)
    np.cumsum(np.array(sizes, dtype=np.int64) ** 2, out=oo[1:])
    return np.concatenate(mats).astype(np.int32), offs, oo


def rec_tuple(r):
    return (int(r["center"]), int(r["u1"]), int(r["u2"]), int(r["pos2"]), int(r[4]), int(r["flags"]), int(r["cluster_size"]), int(r["n_clusters"]))


def test_oracle_matches_second_restatement(orc):
    rng = np.random.default_rng(11)
    tot = flagged = 0
    for t in range(400):
        n = int(rng.integers(2, 36 if t % 8 else 101))
        packed = random_packed(rng, n, t % 3)
        qv = int(rng.integers(0, 2))
        prm = orc.AssignParams(2, 1, int(rng.choice([3000, 3000, 6])), int(rng.choice([50, 50, 2])), 100)
        rec = orc.umi_assign_batch(packed.ravel(), np.array([0, n]), np.array([0, n * n]), prm, np.array([qv], dtype=np.uint8))
        pr = pyref.assign_hier(packed.tolist(), 2, 1, prm.single_threshold, prm.fold_depth, bool(qv))
        for i in range(n):
            a, b = rec[i], pr[i]
            assert (a["center"], a["u1"], a["u2"], a["pos2"], a["off_mean"], bool(a["flags"] & 1), bool(a["flags"] & 2), bool(a["flags"] & 4),
                    a["cluster_size"], a["n_clusters"]) == \
                   (b["center"], b["u1"], b["u2"], b["pos2"], b["off_mean"], b["assigned"], b["skipped"], b["tie_unpin"], b["cluster_size"],
                    b["n_clusters"]), (t, i, a, b)
        tot += n
        flagged += int(rec["flags"][0] & 4 != 0)
    assert tot > 5000 and 0 < flagged < 400


def umi_like_packed(rng, n):
    """a deep job as real data shape it: a few molecules with many reads each, reads 0-2 errors from their molecule, a few chimeric bridges"""
    k = int(rng.integers(3, 12))
    lab = rng.integers(0, k, n)
    err = rng.integers(0, 3, n)
    e = np.where(lab[:, None] == lab[None, :], np.minimum(err[:, None] + err[None, :], 5), np.minimum(3 + rng.integers(0, 3, (n, n)), 5))
    e = np.where(rng.random((n, n)) < 0.01, rng.integers(1, 3, (n, n)), e)
    e = np.triu(e, 1)
    e = e + e.T
    p1, p2 = rng.integers(0, 3, (n, n)), rng.integers(0, 3, (n, n))
    up = e | (0x08000000 << p1) | (0x01000000 << p2)
    lo = e | (0x08000000 << p2.T) | (0x01000000 << p1.T)
    packed = np.where(np.arange(n)[:, None] <= np.arange(n)[None, :], up, lo)
    np.fill_diagonal(packed, 0x10000000 | 0x02000000)
    return packed.astype(np.int32)


def test_myclustering_oracle_matches_second_restatement(orc):
    """ClusterOne_MyClustering.call: the C oracle against the independent Python restatement (oracle/pyref.assign_myclust) on random and
    UMI-shaped matrices of 2 ... 700 reads; the removal path (OneUmiCluster.removeEntries) and the second clusterLocal round must be exercised"""
    rng = np.random.default_rng(77)
    calls = [0]
    orig = pyref.FuIntSet.remove_all

    def counted(self, victims):
        calls[0] += 1
        return orig(self, victims)
    pyref.FuIntSet.remove_all = counted
    try:
        tot = assigned = 0
        for t in range(36):
            n = int(rng.integers(2, 330)) if t % 9 else int(rng.integers(400, 700))
            packed = random_packed(rng, n, t % 3) if t % 2 else umi_like_packed(rng, n)
            qv, fold = int(rng.integers(0, 2)), int(rng.choice([50, 50, 3]))
            rec = orc.umi_assign_batch(packed.ravel(), np.array([0, n]), np.array([0, n * n]), orc.AssignParams(fold_depth=fold, max_hier=0, deep=1),
                                       np.array([qv], dtype=np.uint8))
            pr = pyref.assign_myclust(packed.tolist(), 2, fold, bool(qv))
            for i in range(n):
                a, b = rec[i], pr[i]
                got = (bool(a["flags"] & 1), bool(a["flags"] & 2), bool(a["flags"] & 4), int(a["cluster_size"]), int(a["n_clusters"])) + \
                      ((int(a["center"]), int(a["u1"]), int(a["u2"]), int(a["pos2"]), int(a["off_mean"])) if a["flags"] & 1 else ())
                exp = (b["assigned"], b["skipped"], b["tie_unpin"], b["cluster_size"], b["n_clusters"]) + \
                      ((b["center"], b["u1"], b["u2"], b["pos2"], b["off_mean"]) if b["assigned"] else ())
                assert got == exp, (t, n, i, got, exp)
            assert (rec["flags"] & 8).all()
            tot += n
            assigned += int((rec["flags"] & 1).sum())
    finally:
        pyref.FuIntSet.remove_all = orig
    assert tot > 5000 and assigned > 3000 and calls[0] > 50, (tot, assigned, calls)


def _lingpipe_test_distance():
    """SingleLinkClustererTest$TestDistance (the fixture both of LingPipe's clusterer tests use): A..E"""
    d = np.zeros((5, 5), dtype=np.int64)
    for (a, b), v in {("A", "B"): 1, ("A", "C"): 2, ("A", "D"): 7, ("A", "E"): 5, ("B", "C"): 3, ("B", "D"): 8, ("B", "E"): 6, ("C", "D"): 5,
                      ("C", "E"): 9, ("D", "E"): 4}.items():
        i, j = "ABCDE".index(a), "ABCDE".index(b)
        d[i, j] = d[j, i] = v
    return d


def test_lingpipe_own_unit_test_expectations(orc):
    """CompleteLinkClustererTest.testOne (A!com/aliasi/test/unit/cluster/CompleteLinkClustererTest.class, …java:L53-L123) expects for the fixture
    distances the dendrogram ((A B):1 C):3, (D E):4, root 9: partitionK(2) = {ABC, DE}, partitionK(3) = {AB, C, DE}, partitionK(4) = {AB, C, D, E}.
    The same cuts through partitionDistance: both restatements must produce them (cost <= cut)."""
    d = _lingpipe_test_distance()
    packed = (d | 0x10000000 | 0x02000000).astype(np.int32)
    expect = {1: [{0, 1}], 3: [{0, 1, 2}], 4: [{0, 1, 2}, {3, 4}], 8: [{0, 1, 2}, {3, 4}], 9: [{0, 1, 2, 3, 4}]}
    for cut, clusters in expect.items():
        roots, _ = pyref.complete_link(5, lambda a, b: float(d[a, b]), stop_above=float(cut))
        got = sorted(sorted(c) for r in roots for c in r.partition_distance(float(cut)) if len(c) > 1)
        if cut >= 4:                                         # below 4 D and E have no neighbour and are not part of the clustering at all
            assert got == sorted(sorted(c) for c in clusters), (cut, got)
        rec = orc.umi_assign_batch(packed.ravel(), np.array([0, 5]), np.array([0, 25]), orc.AssignParams(cut, 1, 3000, 50, 100))
        groups = {}
        for i in range(5):
            if rec["center"][i] >= 0:
                groups.setdefault(int(rec["center"][i]), set()).add(i)
        assert sorted(sorted(g) for g in groups.values()) == sorted(sorted(c) for c in clusters), (cut, rec)
    # root cost of the full dendrogram (testOne: "ouch", 9.0)
    roots, _ = pyref.complete_link(5, lambda a, b: float(d[a, b]))
    assert len(roots) == 1 and roots[0].score == 9.0
    # SingleLinkClustererTest: single link on the same fixture merges AB (1), ABC (2), DE (4), all (5)
    for cut, clusters in {1: [{0, 1}], 2: [{0, 1, 2}], 4: [{0, 1, 2}, {3, 4}], 5: [{0, 1, 2, 3, 4}]}.items():
        roots = pyref.single_link(5, lambda a, b: float(d[a, b]), float(cut))
        got = sorted(sorted(r.member_list()) for r in roots if len(r.member_list()) > 1)
        assert got == sorted(sorted(c) for c in clusters), (cut, got)


def test_queue_tie_rule_and_center_rules(orc):
    """equal costs: the pair offered LAST is polled first (BoundedPriorityQueue.java:L458-L464), so of three mutually close reads with one
    far pair the clusterer joins the two highest-numbered ones; two-read clusters take their centre by the quality rule, larger ones the
    least sum of squares with the first in fastutil's iteration order on ties"""
    e = np.array([[0, 1, 1, 5], [1, 0, 1, 5], [1, 1, 0, 1], [5, 5, 1, 0]])     # 0-1-2 a triangle, 3 close to 2 only
    packed = (e | 0x10000000 | 0x02000000).astype(np.int32)
    rec = orc.umi_assign_batch(packed.ravel(), np.array([0, 4]), np.array([0, 16]), orc.AssignParams(1, 1, 3000, 50, 100), np.array([1], dtype=np.uint8))
    # initial pairs in offer order (0,1) (0,2) (0,3) (1,2) (1,3) (2,3): the cheapest offered last is (2,3) -> {2,3}; then (23,x) cost 5; then (1,2)
    # is gone, (0,1) cost 1 -> {0,1}
    assert [int(c >= 0) for c in rec["center"]] == [1, 1, 1, 1]
    assert {frozenset(np.flatnonzero(rec["center"] == c)) for c in set(rec["center"])} == {frozenset({0, 1}), frozenset({2, 3})}
    pr = pyref.assign_hier(packed.tolist(), 1, 1, 3000, 50, True)
    assert [r["center"] for r in pr] == [int(c) for c in rec["center"]]
    assert int(rec["n_clusters"][0]) == 2 and all(int(u) >= 0 for u in rec["u2"])


def _dev_assign(pkg, ctx, mats, offs, oo, params=None, qv=None, n_rows=None, fill=0):
    """slr_umi_assign_dev2 with an arena sized for the jobs above max_hier (slr_umi_assign_deep_job_bytes)"""
    import ctypes as C
    import torch
    L = pkg.gpu_lib()
    d_m = torch.from_numpy(mats).cuda()
    d_o, d_oo = torch.from_numpy(offs).cuda(), torch.from_numpy(oo).cuda()
    m = int(offs[-1]) if n_rows is None else n_rows
    d_rec = torch.full((m, 16), fill, dtype=torch.uint8, device="cuda")
    n_jobs = len(offs) - 1
    max_hier = 100 if params is None else params.max_hier
    deep = sum(int(L.slr_umi_assign_deep_job_bytes(int(n))) for n in np.diff(offs) if n > max_hier)
    nbytes = int(L.slr_umi_assign_scratch_bytes(n_jobs)) + deep
    d_scr = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    d_qv = None if qv is None else torch.from_numpy(qv).cuda()
    st = torch.cuda.current_stream().cuda_stream
    pkg._check(L.slr_umi_assign_dev2(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), n_jobs, m,
                                     C.byref(params) if params is not None else None, None if d_qv is None else d_qv.data_ptr(),
                                     d_scr.data_ptr(), nbytes, d_rec.data_ptr(), st))
    torch.cuda.synchronize()
    return d_rec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)


@pytest.mark.gpu
def test_gpu_assign_matches_oracle_on_arbitrary_matrices(pkg, orc, ctx):
    """slr_umi_assign_dev on injected matrices (cliques, noisy blocks, random graphs; jobs of 1 ... 100 reads and a few above) against the
    oracle: every field of every record, incl. the tie flag"""
    rng = np.random.default_rng(2024)
    for trial, (fold, single_thr) in enumerate(((50, 3000), (2, 3000), (50, 6))):
        mats, offs, oo = make_batch(rng, 1500, max_n=40, big_every=9)
        # two jobs above max_hier: flagged SLR_UA_DEEP, untouched
        n_big = 130
        mats = np.concatenate([mats, random_packed(rng, n_big, 0).ravel()])
        offs = np.append(offs, offs[-1] + n_big)
        oo = np.append(oo, oo[-1] + n_big * n_big)
        qv = rng.integers(0, 2, len(offs) - 1).astype(np.uint8)
        got = _dev_assign(pkg, ctx, mats, offs, oo, pkg.UmiAssignParams(2, 1, single_thr, fold, 100), qv)
        exp = orc.umi_assign_batch(mats, offs, oo, orc.AssignParams(2, 1, single_thr, fold, 100), qv)
        ga, ea = np.frombuffer(got.tobytes(), dtype=np.uint64).reshape(-1, 2), np.frombuffer(exp.tobytes(), dtype=np.uint64).reshape(-1, 2)
        diff = np.flatnonzero((ga != ea).any(axis=1))
        assert len(diff) == 0, (trial, len(diff), diff[:5], got[diff[:3]], exp[diff[:3]])
        assert int((got["flags"] & pkg.UA_ASSIGNED != 0).sum()) > 1000
        assert int((got["flags"] & pkg.UA_DEEP != 0).sum()) == n_big
        if fold == 2:
            assert int((got["flags"] & pkg.UA_SKIPPED != 0).sum()) > 0
        assert 0 < int((got["flags"] & pkg.UA_TIE_UNPIN != 0).sum()) < len(got)


@pytest.mark.gpu
def test_gpu_assign_on_synthetic_umis(pkg, orc, ctx):
    """the fused call on real distance matrices: synthetic (cell, region) jobs (mean 5, up to 100 reads and a few deeper ones) through
    slr_umi_assign, the matrices optionally copied back, and the session entry — records bit-exact against the oracle run on the
    oracle's own matrices; the tie flag is rare on UMI-like data"""
    umis, offs = pkg.synth_umi_jobs(30_000, mean=5.0, cap=160, seed=21)
    u2, o2 = pkg.synth_umi_jobs(60, mean=70.0, cap=160, seed=22)          # some jobs of 33 ... 100 reads and some above
    umis, offs = np.concatenate([umis, u2]), np.concatenate([offs, offs[-1] + o2[1:]])
    assert (np.diff(offs) > 100).sum() >= 5 and ((np.diff(offs) > 32) & (np.diff(offs) <= 100)).sum() >= 5
    em, oo = orc.umi_matrix_batch(umis, offs)
    qv = (np.arange(len(offs) - 1) % 2).astype(np.uint8)
    exp = orc.umi_assign_batch(em, offs, oo, None, qv)
    rec, mat, moo = pkg.cluster_one_hierarchical(ctx, umis, offs, job_qv01=qv, want_matrices=True)
    assert (mat == em).all() and (moo == oo).all()
    assert rec.tobytes() == exp.tobytes()
    rec2 = pkg.cluster_one_hierarchical(ctx, umis, offs, job_qv01=qv)
    assert rec2.tobytes() == exp.tobytes()
    with pkg.UmiSession(ctx, umis, offs) as s:
        rec3 = s.assign(job_qv01=qv)
        assert rec3.tobytes() == exp.tobytes()
        crec = s.cluster(2)                                  # the clusterLocal seam on the same resident matrices still works
        assert crec.tobytes() == orc.umi_cluster_batch(em, offs, oo, 2).tobytes()
    assigned = rec["flags"] & pkg.UA_ASSIGNED != 0
    assert 0.3 < assigned.mean() < 0.95
    assert (rec["flags"] & pkg.UA_DEEP != 0).sum() > 0
    assert (rec["flags"] & pkg.UA_TIE_UNPIN != 0).mean() < 0.02
    # U1 of a centre is the diagonal's distance 0; every assigned read's centre is assigned to itself
    jid = np.repeat(np.arange(len(offs) - 1), np.diff(offs))
    cidx = offs[jid[assigned]] + rec["center"][assigned]
    assert (rec["center"][cidx] == rec["center"][assigned]).all()


@pytest.mark.gpu
def test_gpu_assign_padding_rows_and_base_offset(pkg, orc, ctx):
    """the *_dev entries on buffers 'as slr_umi_dist_dev left them': job offsets that start above 0 and rows behind the last job (ADVICE r1:
    slr_umi_cluster_dev read out of bounds for such rows) — padding rows get empty records, the jobs the oracle's"""
    import torch
    rng = np.random.default_rng(5)
    mats, offs, oo = make_batch(rng, 300, max_n=30, big_every=50)
    base, tail = 37, 91
    m = int(offs[-1])
    offs_b = offs + base
    got = _dev_assign_padded(pkg, ctx, mats, offs_b, oo, m + base + tail)
    exp = orc.umi_assign_batch(mats, offs, oo, None, None)
    assert got[base:base + m].tobytes() == exp.tobytes()
    pad = np.concatenate([got[:base], got[base + m:]])
    assert (pad["center"] == -1).all() and (pad["flags"] == 0).all()
    # the clusterLocal entry on the same padded layout
    d_m, d_o, d_oo = torch.from_numpy(mats).cuda(), torch.from_numpy(offs_b).cuda(), torch.from_numpy(oo).cuda()
    n_rows = m + base + tail
    d_cnt = torch.zeros(n_rows, dtype=torch.int32, device="cuda")
    d_rec = torch.zeros((n_rows, 16), dtype=torch.uint8, device="cuda")
    pkg._check(pkg.gpu_lib().slr_umi_cluster_dev(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), len(offs) - 1, n_rows, 2, None, None,
                                                 d_cnt.data_ptr(), d_rec.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    crec = d_rec.cpu().numpy().view(pkg.UMI_CLUSTER_REC).reshape(-1)
    assert crec[base:base + m].tobytes() == orc.umi_cluster_batch(mats, offs, oo, 2).tobytes()
    assert (crec[:base]["best_key"] == -1).all() and (crec[base + m:]["best_key"] == -1).all()


def _dev_assign_padded(pkg, ctx, mats, offs, oo, n_rows):
    # rec is positional over rows: rec[r] for row r, jobs address rows joff[j] .. joff[j + 1]
    return _dev_assign(pkg, ctx, mats, offs, oo, n_rows=n_rows, fill=0x55)


def _deep_batch(rng, sizes):
    mats, qv = [], []
    for t, n in enumerate(sizes):
        mats.append((random_packed(rng, n, t % 3) if t % 2 else umi_like_packed(rng, n)).ravel())
        qv.append(int(rng.integers(0, 2)))
    sizes = np.array(sizes, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    oo = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
    return np.concatenate(mats).astype(np.int32), offs, oo, np.array(qv, dtype=np.uint8)


@pytest.mark.gpu
def test_gpu_deep_jobs_match_oracle(pkg, orc, ctx):
    """ClusterOne_MyClustering.call on the GPU (umi_assign_deep.cu) against the oracle: jobs of 101 ... 1024 reads on one CTA, larger ones on a
    cluster of 8 CTAs, mixed with small jobs in one batch; every field of every record"""
    rng = np.random.default_rng(909)
    sizes = [int(x) for x in rng.integers(101, 400, 24)] + [1024, 1025, 1500, 2600, 3, 17, 64, 100, 101, 7]
    rng.shuffle(sizes)
    mats, offs, oo, qv = _deep_batch(rng, sizes)
    for fold in (50, 3):
        P = pkg.UmiAssignParams(2, 1, 3000, fold, 100, 1)
        got = _dev_assign(pkg, ctx, mats, offs, oo, P, qv)
        exp = orc.umi_assign_batch(mats, offs, oo, orc.AssignParams(2, 1, 3000, fold, 100, 1), qv)
        bad = np.nonzero(got.tobytes() != exp.tobytes())[0] if False else [i for i in range(len(got)) if rec_tuple(got[i]) != rec_tuple(exp[i])]
        assert not bad, (fold, len(bad), bad[:5], [(rec_tuple(got[i]), rec_tuple(exp[i]), int(np.searchsorted(offs, i, side="right") - 1)) for i in bad[:5]])
        deep = np.repeat(np.diff(offs) > 100, np.diff(offs))
        assert int((got["flags"][deep] & 1).sum()) > 5000 and (got["flags"][deep] & 8).all() and not (got["flags"][~deep] & 8).any()
    # deep = 0 and the arena-less entry point: the large jobs are only flagged
    got0 = _dev_assign(pkg, ctx, mats, offs, oo, pkg.UmiAssignParams(2, 1, 3000, 50, 100, 0), qv)
    assert (got0["flags"][deep] == 8).all() and (got0["center"][deep] == -1).all()


@pytest.mark.gpu
def test_gpu_deep_jobs_through_host_api_and_session(pkg, orc, ctx):
    """slr_umi_assign (host buffers) and the session API size the arena themselves: synthetic UMI batches with jobs above 100 reads"""
    umis, offs = pkg.synth_umi_jobs(300, mean=60.0, cap=900, seed=21)
    assert (np.diff(offs) > 100).sum() >= 20
    rec, mats, oo = pkg.cluster_one_hierarchical(ctx, umis, offs, want_matrices=True)
    exp = orc.umi_assign_batch(mats, offs, oo, orc.AssignParams(), None)
    assert [rec_tuple(r) for r in rec] == [rec_tuple(r) for r in exp]
    deep = np.repeat(np.diff(offs) > 100, np.diff(offs))
    assert int((rec["flags"][deep] & 1).sum()) > 1000
    with pkg.UmiSession(ctx, umis, offs) as s:
        assert [rec_tuple(r) for r in s.assign()] == [rec_tuple(r) for r in exp]


def test_group_by_cell_and_region():
    """host mirror of UmiClustering.groupDataByCellAndRegion + the size filter: groups by (barcode, region), input order inside a group"""
    import __graft_entry__ as g
    pkg = g.load_package()
    rng = np.random.default_rng(3)
    n = 5000
    cell = rng.integers(0, 40, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15 & 0xFFFFFFFF)
    region = rng.integers(0, 30, n).astype(np.int64)
    valid = rng.random(n) < 0.9
    order, off = pkg.group_by_cell_and_region(cell, region, valid)
    seen = set()
    for j in range(len(off) - 1):
        m = order[off[j]:off[j + 1]]
        assert len(m) >= 2 and (np.diff(m) > 0).all() and valid[m].all()
        assert len(set(cell[m].tolist())) == 1 and len(set(region[m].tolist())) == 1
        key = (int(cell[m[0]]), int(region[m[0]]))
        assert key not in seen
        seen.add(key)
    import collections
    cnt = collections.Counter((int(c), int(r)) for c, r, v in zip(cell, region, valid) if v)
    assert seen == {k for k, v in cnt.items() if v >= 2} and off[-1] == sum(v for v in cnt.values() if v >= 2)


def test_fastutil_set_removal_wraparound(orc):
    """IntOpenHashSet + AbstractCollection.removeAll through the set's iterator (OneUmiCluster.removeEntries): the C oracle against the Python
    restatement on sets built to wrap around the table end (home slots in the last few slots, long probe runs), with growth, with key 0, with
    removals that empty runs and removals that trigger the shrink rule — the path the clustering fuzz reaches only a few times"""
    rng = np.random.default_rng(2026)
    wrapped_cases = 0
    for t in range(1500):
        k = int(rng.choice([3, 8, 20, 24, 25, 40, 49, 97, 150, 400]))
        table = 32
        size = 0
        for _ in range(k):                                   # the table size the set will have (for the home-slot filter below)
            size += 1
            if size - 1 >= min(-(-table * 3 // 4), table - 1):
                need, nn = -(-(size + 1) * 4 // 3), 2
                while nn < need:
                    nn *= 2
                table = nn
        pool = np.arange(0, 60000)
        home = np.array([pyref.fastutil_mix(int(x)) & (table - 1) for x in pool[:6000]])
        near_end = pool[:6000][(home >= table - 4) | (home < 2)] if t % 2 else pool[:6000]
        keys = rng.choice(near_end if len(near_end) >= k else pool, size=k, replace=False).astype(np.int32)
        if t % 7 == 0:
            keys[int(rng.integers(k))] = 0
            keys = np.unique(keys)[rng.permutation(len(np.unique(keys)))].astype(np.int32)
        nv = int(rng.integers(1, len(keys) + 1)) if t % 5 else len(keys) - 1
        victims = rng.choice(keys, size=nv, replace=False).astype(np.int32)
        s = pyref.FuIntSet()
        for x in keys:
            s.add(int(x))
        exp_before = s.order()
        s.remove_all([int(v) for v in victims])
        exp_after = s.order()
        got_before, got_after = orc.fu_set_ops(keys, victims)
        assert got_before == exp_before, (t, keys.tolist())
        assert got_after == exp_after, (t, keys.tolist(), victims.tolist())
        assert sorted(exp_after) == sorted(set(keys.tolist()) - set(victims.tolist()))
        wrapped_cases += t % 2
    assert wrapped_cases > 500


def wraparound_job(rng, n):
    """a job of more than 100 reads with ONE cluster whose members sit at the end / start of fastutil's 32-slot table (home slots 28 ... 31, 0, 1) and
    lose a few members to the off-centre removal: hub h (every member within ED 2 of it, so all join its entry), a tight group G that wins the
    centre, outsiders O farther than ED 2 from G — they leave through OneUmiCluster.removeEntries, across the table's wrap-around"""
    home = np.array([pyref.fastutil_mix(int(x)) & 31 for x in range(n)])
    pool = np.nonzero((home >= 28) | (home <= 1))[0]
    pool = pool[pool > 0] if rng.random() < 0.7 else pool                 # with and without key 0
    k = int(rng.integers(6, min(22, len(pool) - 1)))
    mem = rng.choice(pool, size=k + 1, replace=False)
    h, rest = int(mem[0]), mem[1:]
    n_out = int(rng.integers(1, 4))
    O, G = rest[:n_out], rest[n_out:]
    e = np.full((n, n), 5, dtype=np.int64)
    for g in G:
        e[h, g] = e[g, h] = 2
        for g2 in G:
            e[g, g2] = int(rng.integers(0, 2))
    for o in O:
        e[h, o] = e[o, h] = 2
        for g in G:
            e[o, g] = e[g, o] = 4
    e = np.triu(e, 1)
    e = e + e.T
    p1, p2 = rng.integers(0, 3, (n, n)), rng.integers(0, 3, (n, n))
    up = e | (0x08000000 << p1) | (0x01000000 << p2)
    lo = e | (0x08000000 << p2.T) | (0x01000000 << p1.T)
    packed = np.where(np.arange(n)[:, None] <= np.arange(n)[None, :], up, lo)
    np.fill_diagonal(packed, 0x10000000 | 0x02000000)
    return packed.astype(np.int32)


def test_myclustering_removal_across_the_table_end(orc):
    """oracle vs the Python restatement on jobs built so that the off-centre removal walks fastutil's wrap-around (wrapped entries, removal through
    the set's own remove) — counted, so the path is known to run"""
    rng = np.random.default_rng(8)
    cnt = {"wrapped": 0}
    orig = pyref.FuIntSet._shift

    def shift(self, pos, wrapped=None):
        n0 = len(wrapped) if wrapped is not None else 0
        orig(self, pos, wrapped)
        if wrapped is not None:
            cnt["wrapped"] += len(wrapped) - n0
    pyref.FuIntSet._shift = shift
    try:
        removed_jobs = 0
        for t in range(80):
            n = int(rng.integers(101, 330))
            packed = wraparound_job(rng, n)
            qv = int(rng.integers(0, 2))
            rec = orc.umi_assign_batch(packed.ravel(), np.array([0, n]), np.array([0, n * n]), orc.AssignParams(), np.array([qv], dtype=np.uint8))
            pr = pyref.assign_myclust(packed.tolist(), 2, 50, bool(qv))
            for i in range(n):
                a, b = rec[i], pr[i]
                assert (bool(a["flags"] & 1), int(a["cluster_size"]), int(a["n_clusters"])) == (b["assigned"], b["cluster_size"], b["n_clusters"]), (t, i)
                if b["assigned"]:
                    assert (int(a["center"]), int(a["u1"]), int(a["u2"]), int(a["pos2"]), int(a["off_mean"])) == \
                           (b["center"], b["u1"], b["u2"], b["pos2"], b["off_mean"]), (t, i)
            removed_jobs += int((rec["flags"] & 1).sum() > 0 and int(rec["cluster_size"].max()) < int((packed.ravel() & 0xFF <= 2).reshape(n, n).sum(axis=1).max()))
    finally:
        pyref.FuIntSet._shift = orig
    assert cnt["wrapped"] >= 10 and removed_jobs >= 40, (cnt, removed_jobs)


@pytest.mark.gpu
def test_gpu_deep_removal_across_the_table_end(pkg, orc, ctx):
    """the same constructed jobs through the large-job kernels: thread-per-cluster local tables (<= 24 members) with wrap-around removal"""
    rng = np.random.default_rng(8)
    sizes = [int(rng.integers(101, 330)) for _ in range(80)]
    mats = np.concatenate([wraparound_job(rng, n).ravel() for n in sizes]).astype(np.int32)
    sizes = np.array(sizes, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    oo = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
    qv = (np.arange(len(sizes)) % 2).astype(np.uint8)
    got = _dev_assign(pkg, ctx, mats, offs, oo, None, qv)
    exp = orc.umi_assign_batch(mats, offs, oo, None, qv)
    assert [rec_tuple(r) for r in got] == [rec_tuple(r) for r in exp]
    assert int((got["flags"] & 1).sum()) > 400
