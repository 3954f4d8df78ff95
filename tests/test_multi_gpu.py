"""Several GPUs behind one caller (slr_multi_*): the C-ABI route of the single-JVM reference to all devices of a box.
Runs on however many devices are visible (1 on the default test box; `gpurun --gpus 2` exercises the real split, the peer-access counter
reduction and the job dealing)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_multi_matches_single_device_and_oracle(pkg, orc, ctx):
    import torch
    n_dev = torch.cuda.device_count()
    mg = pkg.MultiGpu(0)
    assert mg.n_devices == n_dev
    wl = pkg.synth_whitelist(200_000, 99)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    n = 300_001                                              # not a multiple of the device count
    slices, anchor, _ = pkg.synth_reads(wl, n, seed=31)
    mg.load_barcodes(wl, rank)
    got = mg.assign_barcodes(slices, anchor, 2)
    table = pkg.BarcodesMapForBCfinding(ctx, wl, rank)
    one = pkg.Parser(ctx, table, 2).assign_barcodes(slices, anchor)
    assert (got == one).all()
    sel = np.arange(0, n, 97)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices[sel], anchor[sel], 2)
    assert (got[sel] == exp).all()
    # counters: summed over the replicas on the device = the single-device counters = the histogram of the records
    cm, c1 = mg.counts(), table.counts()
    assert (cm == c1).all()
    ok = (got["flags"] & 1) != 0
    assert cm.sum() == ok.sum() and (cm.sum(axis=0) == np.bincount(got["ed"][ok], minlength=3)).all()
    mg.assign_barcodes(slices, anchor, 2)                    # a second batch accumulates on every replica
    assert (mg.counts() == 2 * c1).all()
    mg.reset_counts()
    assert mg.counts().sum() == 0
    # UMI seams: whole jobs dealt to the devices
    umis, offs = pkg.synth_umi_jobs(40_000, mean=5.0, cap=300, seed=8)
    em, oo = orc.umi_matrix_batch(umis, offs)
    qv = (np.arange(len(offs) - 1) % 3 == 0).astype(np.uint8)
    assert mg.umi_assign(umis, offs, job_qv01=qv).tobytes() == orc.umi_assign_batch(em, offs, oo, None, qv).tobytes()
    assert mg.umi_cluster(umis, offs, 2).tobytes() == orc.umi_cluster_batch(em, offs, oo, 2).tobytes()
    m, moo = mg.umi_dist(umis, offs)
    assert (m == em).all() and (moo == oo).all()
    # jobs above 100 reads (ClusterOne_MyClustering on every device: each shard sizes its own arena)
    du, doff = pkg.synth_umi_jobs(60, mean=150.0, cap=700, seed=12)
    for cap_, seed_ in ((1500, 31), (4200, 32)):            # + a cluster-of-8-CTAs job and a cooperative-grid job: they end up on the LAST device,
        gu, go = pkg.synth_umi_jobs(1, mean=1e9, cap=cap_, seed=seed_)      # whose function attributes (192 KB of dynamic shared memory) are its own
        du, doff = np.concatenate([du, gu]), np.concatenate([doff, doff[-1] + go[1:]]).astype(np.int64)
    dm, doo = orc.umi_matrix_batch(du, doff)
    drec = mg.umi_assign(du, doff)
    assert drec.tobytes() == orc.umi_assign_batch(dm, doff, doo).tobytes()
    assert int(((drec["flags"] & 9) == 9).sum()) > 1000
    if n_dev > 1:
        assert mg.peer_access                                # NVSwitch box: the counter reduction reads the peers' HBM directly
    mg.close()


@pytest.mark.gpu
def test_multi_rejects_foreign_table_and_bad_devices(pkg, ctx):
    import ctypes as C
    h = C.c_void_p()
    ids = np.array([4711], dtype=np.int32)
    assert pkg.gpu_lib().slr_multi_create(1, ids.ctypes.data, 2, C.byref(h)) == pkg.SLR_E_INVALID
    a, b = pkg.MultiGpu(1), pkg.MultiGpu(1)
    wl = pkg.synth_whitelist(1000, 1)
    a.load_barcodes(wl)
    sl, an, _ = pkg.synth_reads(wl, 10, seed=1)
    out = np.empty(10, dtype=pkg.BC_RESULT)
    rc = pkg.gpu_lib().slr_multi_bc_assign(b.h, a.table, 1, 2, 1, sl.ctypes.data, 32, 32, None, an.ctypes.data, 10, out.ctypes.data)
    assert rc == pkg.SLR_E_INVALID
    # a single-device table handed to another context is refused too
    other = pkg.Context(0)
    t = pkg.BarcodesMapForBCfinding(ctx, wl)
    rc = pkg.gpu_lib().slr_bc_assign(other.h, t.h, 1, 2, 1, sl.ctypes.data, 32, 32, None, an.ctypes.data, 10, out.ctypes.data)
    assert rc == pkg.SLR_E_INVALID and b"another context" in pkg.gpu_lib().slr_last_error()
