"""N > 1 path on the CPU (gloo, world_size 2): reads shard by rank with no data-path collective; the only exchange is the
additive per-barcode x ED counter table (BarcodesAssigned.tsv), all-reduced exactly like bench.py does over NCCL.
The per-shard compute here is the CPU oracle (test infrastructure) — the point is the sharding + merge logic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_per_rank, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import orc
    wl = pkg.synth_whitelist(20000, 3)
    slices, anchor, _ = pkg.synth_reads(wl, n_per_rank, seed=9, first=rank * n_per_rank)     # this rank's shard
    res, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, np.arange(1, len(wl) + 1, dtype=np.int32)), slices, anchor, 1, n_threads=2)
    counts = np.zeros((len(wl), 3), dtype=np.int64)
    ok = (res["flags"] & 1) == 1
    np.add.at(counts, (res["rank"][ok] - 1, res["ed"][ok]), 1)
    t = torch.from_numpy(counts)
    dist.all_reduce(t)                                    # cross-shard merge
    gathered = [None] * world
    dist.all_gather_object(gathered, res.tobytes())
    if rank == 0:
        q.put((t.numpy().copy(), b"".join(gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_read_sharding_and_counter_merge():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    pkg.build()
    from oracle import orc
    orc.build()
    world, n_per_rank = 2, 1500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, blob = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run over the whole read range
    wl = pkg.synth_whitelist(20000, 3)
    slices, anchor, _ = pkg.synth_reads(wl, world * n_per_rank, seed=9)
    res, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, np.arange(1, len(wl) + 1, dtype=np.int32)), slices, anchor, 1)
    assert res.tobytes() == blob                           # positional results: concatenation of the shards
    exp = np.zeros((len(wl), 3), dtype=np.int64)
    ok = (res["flags"] & 1) == 1
    np.add.at(exp, (res["rank"][ok] - 1, res["ed"][ok]), 1)
    assert (counts == exp).all() and counts.sum() == ok.sum() > 0


# ---- cross-shard UMI merge (north star: "NCCL ... only for the cross-shard UMI-cluster merge"); gloo on the CPU ------------
def _umi_worker(rank, world, port, cuts, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    import workloads
    pkg = g.load_package()
    from oracle import orc
    umis, offs = workloads.umi_jobs(77, 12, n_jobs=40, max_n=30)           # the global, (cell, region)-sorted stream
    lo, hi = cuts[rank], cuts[rank + 1]
    # local view: the jobs that intersect [lo, hi), clipped; key = (global job index, 0)
    j0 = int(np.searchsorted(offs, lo, side="right")) - 1
    j1 = int(np.searchsorted(offs, hi, side="left"))
    local = np.clip(offs[j0:j1 + 1], lo, hi) - lo if hi > lo else np.zeros(1, dtype=np.int64)
    cap = 64
    buf = torch.zeros((hi - lo + cap, 16), dtype=torch.uint8)
    buf[:hi - lo] = torch.from_numpy(umis[lo:hi])
    row0, n_rows, moffs = pkg.UmiShardMerger(cap=cap).merge(buf, hi - lo, local, (j0, 0), (j1 - 1, 0))
    mine = buf[row0:row0 + n_rows].numpy()
    mats, oo = orc.umi_matrix_batch(mine, moffs, 12, n_threads=1)
    first_job = j0 + (1 if row0 else 0)
    crec = orc.umi_cluster_batch(mats, moffs, oo, 2, n_threads=1)          # clusterLocal needs whole jobs too (neighbour sets span the job)
    out = {first_job + i: (mats[oo[i]:oo[i + 1]].copy(), crec[moffs[i]:moffs[i + 1]].tobytes())
           for i in range(len(moffs) - 1) if moffs[i + 1] > moffs[i]}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_umi_cross_shard_merge():
    import pytest
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    import workloads
    pkg = g.load_package()
    from oracle import orc
    orc.build()
    umis, offs = workloads.umi_jobs(77, 12, n_jobs=40, max_n=30)
    exp, eo = orc.umi_matrix_batch(umis, offs, 12)
    m = int(offs[-1])
    # shard cuts: inside a job, exactly on a job boundary, a shard lying completely inside one job, an empty shard
    mid = int(offs[20] + 3)
    cuts = [0, int(offs[7] + 2), int(offs[12]), mid, mid + 2, mid + 2, m]
    assert offs[20] < mid + 2 < offs[21] and offs[21] - offs[20] >= 8
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_umi_worker, args=(r, world, port, cuts, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    gathered = None
    for _ in range(150):                                      # poll: a crashed worker fails the test at once
        try:
            gathered = q.get(timeout=2)
            break
        except _queue.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), [p.exitcode for p in procs]
    assert gathered is not None
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = {}
    for d in gathered:
        for j, mat in d.items():
            assert j not in seen                              # every job is clustered on exactly one rank
            seen[j] = mat
    assert sorted(seen) == list(range(len(offs) - 1))
    ecl = orc.umi_cluster_batch(exp, offs, eo, 2)
    for j, (mat, crec) in seen.items():
        assert (mat == exp[eo[j]:eo[j + 1]]).all(), j         # ... as ONE job: identical to the unsharded matrix
        assert crec == ecl[offs[j]:offs[j + 1]].tobytes(), j  # ... and so are its neighbour sets / chosen entries
    # the pure planning function: chains and the size guard
    P = pkg.UmiShardMerger.plan
    meta = [[0, 0, 5, 3, 0, 4], [3, 0, 2, 3, 0, 1], [3, 0, 4, 9, 0, 3]]       # job 3 spans ranks 0, 1 (entirely) and 2
    assert P(meta, 0, 64) == (0, [(1, 2), (2, 4)]) and P(meta, 1, 64) == (2, []) and P(meta, 2, 64) == (4, [])
    with pytest.raises(pkg.SiceloreGpuError):
        P(meta, 0, 5)


def _guided_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import orc
    w = pkg.synth_guided(1200, 12, seed=21, n_groups=200, group_size=6, pm=2, post_len=6)          # every rank derives the same batch ...
    lo, hi = rank * 600, (rank + 1) * 600                                                          # ... and takes its contiguous shard of the reads
    res, _, _ = orc.guided_batch(w["group_keys"], w["group_offsets"], w["slices"][lo:hi], w["anchor"][lo:hi], w["group_id"][lo:hi],
                                 (np.arange(1200) % 3)[lo:hi].astype(np.int32), 12, 2, 6, n_threads=2)             # candidate sets replicated
    found = torch.tensor([int((res["n_distinct"] > 0).sum()), int((res["n_distinct"] > 1).sum())])
    dist.all_reduce(found)                                                                          # run statistics are the only exchange
    gathered = [None] * world
    dist.all_gather_object(gathered, res.tobytes())
    if rank == 0:
        q.put((found.numpy().copy(), b"".join(gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_guided_search_shards_by_read():
    """Illumina-guided search at N > 1: reads shard by rank, the candidate sets are replicated, results are positional — no data-path
    collective (the per-shard compute here is the CPU oracle; the GPU path is checked against it elsewhere)"""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    pkg.build()
    from oracle import orc
    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_guided_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    found, blob = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = pkg.synth_guided(1200, 12, seed=21, n_groups=200, group_size=6, pm=2, post_len=6)
    res, _, _ = orc.guided_batch(w["group_keys"], w["group_offsets"], w["slices"], w["anchor"], w["group_id"], (np.arange(1200) % 3).astype(np.int32), 12, 2, 6)
    assert res.tobytes() == blob
    assert found[0] == (res["n_distinct"] > 0).sum() > 600 and found[1] == (res["n_distinct"] > 1).sum()
