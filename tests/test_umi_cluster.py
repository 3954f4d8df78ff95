"""Neighbour-set clustering seam: ClusterOne_MyClustering.clusterLocal
(F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class, ClusterOne_MyClustering.java:L175-L219).

CPU: the C oracle against the stream-by-stream Python restatement.  GPU: slr_umi_cluster through the C ABI against the
oracle on the same seeded UMIs (bit-exact records, bit-exact matrices)."""
import numpy as np
import pytest

import workloads
from oracle import pyref


def _random_matrix(rng, n, symmetric=True, p_close=0.25):
    """packed ints with ED 0..5 in the low byte and arbitrary shift flags above (getED must ignore them)."""
    ed = np.where(rng.random((n, n)) < p_close, rng.integers(0, 3, (n, n)), rng.integers(3, 6, (n, n))).astype(np.int32)
    if symmetric:
        ed = np.triu(ed, 1)
        ed = ed + ed.T
    np.fill_diagonal(ed, 0)
    flags = (rng.integers(0, 64, (n, n)).astype(np.int32)) << 24
    return ed | flags


def _clusters(rec, n):
    groups = {}
    for c in range(n):
        k = int(rec["best_key"][c])
        if k >= 0:
            groups.setdefault(k, set()).add(c)
    return {frozenset(v) for v in groups.values()} or None


def _rank_of(order, n):
    rank = np.full(n, 2 ** 30, dtype=np.int32)
    for i, k in enumerate(order):
        rank[k] = i
    return rank


ORDERS = {
    "ascending": None,
    "fastutil": pyref.fastutil_key_order,
    "reversed": lambda keys: sorted(keys, reverse=True),
}


@pytest.mark.parametrize("order_name", list(ORDERS))
@pytest.mark.parametrize("ed", [0, 1, 2, 4])
def test_oracle_matches_python_restatement(orc, order_name, ed):
    rng = np.random.default_rng(100 + ed)
    key_order = ORDERS[order_name]
    for trial in range(40):
        n = int(rng.integers(1, 45))
        m = _random_matrix(rng, n, symmetric=trial % 3 != 0, p_close=float(rng.choice([0.05, 0.3, 0.8])))
        member = (rng.random(n) < 0.7).astype(np.uint8) if trial % 2 else None
        indices = [i for i in range(n) if member is None or member[i]]
        want = pyref.cluster_local(m.tolist(), indices, ed, key_order)
        rank = None
        if key_order is not None:
            rank = _rank_of(key_order(indices), n)     # any superset of the keys in a consistent order will do
            # the order must be the one of the KEYS only: recompute it from the keys the oracle finds
            rec0 = orc.umi_cluster_batch(m.ravel(), [0, n], [0, n * n], ed, member)
            keys = [i for i in range(n) if rec0["n_neighbours"][i] > 1]
            rank = _rank_of(key_order(keys), n)
        rec = orc.umi_cluster_batch(m.ravel(), [0, n], [0, n * n], ed, member, rank)
        assert _clusters(rec, n) == want, (trial, n)
        # record fields
        for c in range(n):
            nb = [v for v in indices if pyref._i8(int(m[c, v]) & 0xFFFFFF) <= ed] if c in indices else []
            assert rec["n_neighbours"][c] == len(nb)
            if len(nb) <= 1:
                assert rec["best_key"][c] == -1 and rec["n_ties"][c] == 0
            else:
                k = int(rec["best_key"][c])
                assert rec["n_neighbours"][k] == rec["best_count"][c]
                tied = [l for l in indices if rec["n_neighbours"][l] == rec["best_count"][c] and rec["n_neighbours"][l] > 1
                        and pyref._i8(int(m[l, c]) & 0xFFFFFF) <= ed]
                assert rec["n_ties"][c] == len(tied) and k in tied


def test_no_keys_is_optional_empty(orc):
    m = np.full((5, 5), 5, dtype=np.int32)
    np.fill_diagonal(m, 0)
    rec = orc.umi_cluster_batch(m.ravel(), [0, 5], [0, 25], 2)
    assert (rec["best_key"] == -1).all() and (rec["n_neighbours"] == 1).all()
    assert pyref.cluster_local(m.tolist(), range(5), 2) is None


def test_first_maximum_wins_and_ties_are_counted(orc):
    # 0-1-2 chain at ED 1: N(0) = {0,1}, N(1) = {0,1,2}, N(2) = {1,2}; 3, 4 a separate pair with equal counts
    m = np.full((5, 5), 5, dtype=np.int32)
    np.fill_diagonal(m, 0)
    for a, b in ((0, 1), (1, 2), (3, 4)):
        m[a, b] = m[b, a] = 1
    rec = orc.umi_cluster_batch(m.ravel(), [0, 5], [0, 25], 1)
    assert rec["best_key"].tolist() == [1, 1, 1, 3, 3] and rec["n_ties"].tolist() == [1, 1, 1, 2, 2]
    rec = orc.umi_cluster_batch(m.ravel(), [0, 5], [0, 25], 1, rank=np.array([4, 3, 2, 1, 0], dtype=np.int32))
    assert rec["best_key"].tolist() == [1, 1, 1, 4, 4]


def test_oracle_on_real_matrices_match_python(orc):
    umis, offs = workloads.umi_jobs(21, 12, n_jobs=30, max_n=30)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    rec = orc.umi_cluster_batch(mats, offs, oo, 2)
    for j in range(len(offs) - 1):
        a, n = int(offs[j]), int(offs[j + 1] - offs[j])
        m = mats[oo[j]:oo[j + 1]].reshape(n, n)
        want = pyref.cluster_local(m.tolist(), range(n), 2)
        assert _clusters(rec[a:a + n], n) == want


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("ed", [0, 1, 2, 3])
def test_gpu_cluster_matches_oracle(pkg, orc, ctx, ed):
    umis, offs = workloads.umi_jobs(300 + ed, 12, n_jobs=120, max_n=70)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    want = orc.umi_cluster_batch(mats, offs, oo, ed)
    rec, got_m, got_oo = pkg.cluster_local(ctx, umis, offs, ed, want_matrices=True)
    assert np.array_equal(got_oo, oo) and np.array_equal(got_m, mats)
    assert rec.tobytes() == want.tobytes()
    assert pkg.clusters_from_records(rec, offs) == pkg.clusters_from_records(want, offs)
    # without the matrix copy-back
    rec2 = pkg.cluster_local(ctx, umis, offs, ed)
    assert rec2.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_gpu_cluster_member_subset_and_rank(pkg, orc, ctx):
    rng = np.random.default_rng(5)
    umis, offs = workloads.umi_jobs(77, 12, n_jobs=150, max_n=60)
    m = int(offs[-1])
    member = (rng.random(m) < 0.6).astype(np.uint8)
    rank = np.concatenate([rng.permutation(int(offs[j + 1] - offs[j])) for j in range(len(offs) - 1)]).astype(np.int32)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    for mem, rk in ((member, None), (None, rank), (member, rank)):
        want = orc.umi_cluster_batch(mats, offs, oo, 2, mem, rk)
        got = pkg.cluster_local(ctx, umis, offs, 2, member=mem, rank=rk)
        assert got.tobytes() == want.tobytes()
    assert (want["n_ties"] > 1).any()          # the workload does exercise the tie rule


@pytest.mark.gpu
def test_gpu_cluster_deep_job_between_small_ones_with_member_and_rank(pkg, orc, ctx):
    """the deep kernels (jobs of >= 1024 reads) with a member subset and a caller order, a deep job that starts and ends inside a
    32-read tile shared with small jobs, and two deep jobs in one launch"""
    rng = np.random.default_rng(11)
    small, so = workloads.umi_jobs(5, 12, n_jobs=9, max_n=20)
    deep1, _ = pkg.synth_umi_jobs(1, mean=1e9, cap=1500, seed=3)
    deep2, _ = pkg.synth_umi_jobs(1, mean=1e9, cap=1100, seed=4)
    k = int(so[4])
    umis = np.concatenate([small[:k], deep1, small[k:], deep2])
    sizes = np.concatenate([np.diff(so[:5]), [len(deep1)], np.diff(so[4:]), [len(deep2)]])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    assert int(offs[4]) % 32 != 0 and int(offs[5]) % 32 != 0 and len(umis) == offs[-1] and offs[5] - offs[4] == len(deep1)
    m = len(umis)
    member = (rng.random(m) < 0.8).astype(np.uint8)
    rank = np.concatenate([rng.permutation(int(n)) for n in sizes]).astype(np.int32)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    for ed in (1, 2):
        for mem, rk in ((None, None), (member, None), (None, rank), (member, rank)):
            want = orc.umi_cluster_batch(mats, offs, oo, ed, mem, rk)
            got = pkg.cluster_local(ctx, umis, offs, ed, member=mem, rank=rk)
            assert got.tobytes() == want.tobytes(), (ed, mem is not None, rk is not None)
    assert (want["best_key"][offs[4]:offs[5]] >= 0).sum() > 500


@pytest.mark.gpu
def test_gpu_cluster_deep_job_and_many_ranges(pkg, orc, ctx):
    # one deep job (row loop over thousands of reads) and a batch large enough to be cut into several ranges
    umis, offs = pkg.synth_umi_jobs(1, mean=1e9, cap=3000, seed=8)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    want = orc.umi_cluster_batch(mats, offs, oo, 2)
    got = pkg.cluster_local(ctx, umis, offs, 2)
    assert got.tobytes() == want.tobytes()
    umis, offs = pkg.synth_umi_jobs(400000, mean=4.0, cap=2000, seed=4)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    assert int(oo[-1]) > (1 << 23)             # more than one range of the ping-pong loop
    want = orc.umi_cluster_batch(mats, offs, oo, 2)
    got = pkg.cluster_local(ctx, umis, offs, 2)
    assert got.tobytes() == want.tobytes()
    # every key lands in exactly one cluster and clusters never cross jobs
    cl = pkg.clusters_from_records(got, offs)
    for j, sets in enumerate(cl[:2000]):
        n = int(offs[j + 1] - offs[j])
        seen = [x for s in sets for x in s]
        assert len(seen) == len(set(seen)) and all(0 <= x < n for x in seen)


@pytest.mark.gpu
def test_gpu_cluster_dev_entry_and_refusals(pkg, orc, ctx):
    import torch
    umis, offs = workloads.umi_jobs(9, 12, n_jobs=40, max_n=40)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    want = orc.umi_cluster_batch(mats, offs, oo, 1)
    m = int(offs[-1])
    d_m = torch.from_numpy(mats).cuda()
    d_jo, d_oo = torch.from_numpy(offs).cuda(), torch.from_numpy(oo).cuda()
    d_cnt = torch.empty(m, dtype=torch.int32, device="cuda")
    d_rec = torch.empty(m * 4, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = pkg.gpu_lib().slr_umi_cluster_dev(ctx.h, d_m.data_ptr(), d_jo.data_ptr(), d_oo.data_ptr(), len(offs) - 1, m, 1, None, None,
                                           d_cnt.data_ptr(), d_rec.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    assert d_rec.cpu().numpy().tobytes() == want.tobytes()
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.cluster_local(ctx, umis, offs, 6)
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.cluster_local(ctx, umis, offs, 2, umi_len=15)
    assert len(pkg.cluster_local(ctx, np.zeros((0, 16), np.uint8), np.zeros(1, np.int64), 2)) == 0


@pytest.mark.gpu
def test_gpu_session_two_call_protocol(pkg, orc, ctx):
    """create -> cluster(rank = None) -> the caller derives the iteration rank of the keys -> cluster(rank) -> cluster(member, rank)
    -> matrices, all on the matrices of ONE run of the distance kernels"""
    rng = np.random.default_rng(3)
    umis, offs = workloads.umi_jobs(123, 12, n_jobs=200, max_n=70)
    deep, _ = pkg.synth_umi_jobs(1, mean=1e9, cap=1300, seed=6)
    umis = np.concatenate([umis, deep])
    offs = np.concatenate([offs, [offs[-1] + len(deep)]]).astype(np.int64)
    mats, oo = orc.umi_matrix_batch(umis, offs, 12)
    l0 = pkg.launch_count()
    with pkg.UmiSession(ctx, umis, offs) as s:
        first = s.cluster(2)
        assert first.tobytes() == orc.umi_cluster_batch(mats, offs, oo, 2).tobytes()
        rank = np.full(len(umis), 2 ** 30, dtype=np.int32)             # the caller's map: here the restated fastutil order of the keys
        for j in range(len(offs) - 1):
            a, b = int(offs[j]), int(offs[j + 1])
            keys = [i for i in range(b - a) if first["n_neighbours"][a + i] > 1]
            for r, k in enumerate(pyref.fastutil_key_order(keys)):
                rank[a + k] = r
        second = s.cluster(2, rank=rank)
        want = orc.umi_cluster_batch(mats, offs, oo, 2, None, rank)
        assert second.tobytes() == want.tobytes()
        assert (second["best_key"] != first["best_key"]).any()         # the order did change some choices
        member = (rng.random(len(umis)) < 0.5).astype(np.uint8)
        third = s.cluster(2, member=member, rank=rank)
        assert third.tobytes() == orc.umi_cluster_batch(mats, offs, oo, 2, member, rank).tobytes()
        got_m, got_oo = s.matrices()
        assert np.array_equal(got_m, mats) and np.array_equal(got_oo, oo)
    # distance kernels once (3 + 3 launches), cluster kernels three times (4 each)
    assert pkg.launch_count() - l0 == 6 + 3 * 4
    with pkg.UmiSession(ctx, np.zeros((0, 16), np.uint8), np.zeros(1, np.int64)) as s:
        assert len(s.cluster(2)) == 0 and len(s.matrices()[0]) == 0
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.UmiSession(ctx, umis, offs[::-1].copy())
