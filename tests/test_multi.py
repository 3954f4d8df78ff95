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
