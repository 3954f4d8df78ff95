"""CPU tests of the GPU algorithm: tests/host_sim replays the kernels' per-lane code (bc_core.cuh, umi_core.cuh,
slr_table.cuh and the real table builder) lane by lane on the CPU.  Everything except the warp intrinsics of
bc_assign.cu / umi_dist.cu is therefore checked against the oracle without a GPU."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import workloads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_sim(sim, orc, wl, rank, slices, anchors, ed, pm, three_prime, lens=None):
    got = np.zeros(len(slices), dtype=orc.BC_RESULT)
    counts = np.zeros(len(wl) * 3, dtype=np.uint64)
    loads = C.c_longlong(0)
    st = np.zeros(4, dtype=np.int64)
    sim.sim_bc_assign(wl.ctypes.data, None if rank is None else rank.ctypes.data, len(wl), 0, ed, pm, int(three_prime),
                      slices.ctypes.data, slices.shape[1], min(32, slices.shape[1]), None if lens is None else lens.ctypes.data,
                      anchors.ctypes.data, len(slices), got.ctypes.data, counts.ctypes.data, C.byref(loads), st.ctypes.data)
    return got, counts.reshape(-1, 3), loads.value, st


@pytest.mark.parametrize("three_prime", [True, False])
@pytest.mark.parametrize("ed", [0, 1, 2])
@pytest.mark.parametrize("skew", [False, True])
def test_sim_vs_oracle_adversarial(sim, orc, three_prime, ed, skew):
    _, slices, anchors, wl = workloads.adversarial(1000 + ed + 10 * skew, three_prime, 250, skew=skew)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, three_prime)
    got, counts, _, _ = run_sim(sim, orc, wl, rank, slices, anchors, ed, 2, three_prime)
    assert (got == exp).all()
    assert counts.sum() == (exp["flags"] & 1).sum()


@pytest.mark.parametrize("three_prime", [True, False])
@pytest.mark.parametrize("ed", [1, 2])
def test_sim_vs_oracle_dense_overflow(sim, orc, three_prime, ed):
    """whole digit-group clusters in the list: buckets overflow into the stash, many slots pass the tag filter"""
    _, slices, anchors, wl = workloads.adversarial(2000 + ed, three_prime, 60, skew=True, dense=True, nrand=50)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, three_prime)
    got, _, _, stash = run_sim(sim, orc, wl, rank, slices, anchors, ed, 2, three_prime)
    assert stash.sum() > 100
    assert (got == exp).all()


@pytest.mark.parametrize("pm", [0, 1, 3, 4])
def test_sim_plusminus(sim, orc, pm):
    _, slices, anchors, wl = workloads.adversarial(3000 + pm, True, 120, anchor=10)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), slices, anchors, 2, pm, True)
    got, _, _, _ = run_sim(sim, orc, wl, None, slices, anchors, 2, pm, True)
    assert (got == exp).all()
    assert ((exp["flags"] & orc.F_EXCEPTION) == 0).all()          # 3' span [anchor-pm-4, anchor+pm+16) fits for anchor = 10


def test_sim_ragged_lens_and_empty(sim, orc):
    _, slices, anchors, wl = workloads.adversarial(77, True, 64)
    lens = np.random.default_rng(1).integers(0, 33, size=64).astype(np.int32)
    exp = np.zeros(64, dtype=orc.BC_RESULT)
    bs = orc.BarcodeSet(wl)
    for i in range(64):
        r, _ = orc.assign_barcode_batch(bs, slices[i:i + 1], anchors[i:i + 1], 2, 2, True, slice_len=int(lens[i]))
        exp[i] = r[0]
    got, _, _, _ = run_sim(sim, orc, wl, None, slices, anchors, 2, 2, True, lens=lens)
    assert (got == exp).all()
    got0, _, _, _ = run_sim(sim, orc, wl, None, slices[:0], anchors[:0], 2, 2, True)
    assert len(got0) == 0


def test_sim_synthetic_737k(sim, orc, pkg):
    """configs[1]/[2]-shaped input at a size the oracle finishes in seconds"""
    wl = pkg.synth_whitelist(737280, 737)
    slices, anchors, truth = pkg.synth_reads(wl, 3000, seed=1)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    bs = orc.BarcodeSet(wl, rank)
    for ed in (1, 2):
        exp, probes = orc.assign_barcode_batch(bs, slices, anchors, ed, 2, True)
        got, _, loads, _ = run_sim(sim, orc, wl, rank, slices, anchors, ed, 2, True)
        assert (got == exp).all()
        assert loads * 5 < probes                      # bucket loads vs reference probes
    ok = (exp["flags"] & 1) == 1
    assert ok.mean() > 0.6
    hit = truth[ok] >= 0
    assert (exp["bc"][ok][hit] == wl[truth[ok][hit]]).mean() > 0.9       # README.md:100,180: ED 2 trades accuracy for yield


def test_sim_lazy_level2_is_exact(sim, orc, pkg):
    """slr_level2_plan skips ED-2 searches that cannot reach the record: same bytes as running all of them (what the
    reference does), on adversarial inputs (many planted ED <= 1 / ED 2 neighbours, HashSet ties) and on synthetic
    reads, with fewer bucket loads on the latter."""
    sim.sim_set_force_all_l2.argtypes = [C.c_int]
    cases = [workloads.adversarial(4100 + i, tp, 300, skew=bool(i & 1))[1:] for i, tp in enumerate((True, False, True, False))]
    wl = pkg.synth_whitelist(200000, 9)
    sl, an, _ = pkg.synth_reads(wl, 3000, seed=3)
    cases.append((sl, an, wl))
    for i, (slices, anchors, wl) in enumerate(cases):
        exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), slices, anchors, 2, 2, i % 2 == 0)
        try:
            sim.sim_set_force_all_l2(1)
            full, _, loads_full, _ = run_sim(sim, orc, wl, None, slices, anchors, 2, 2, i % 2 == 0)
        finally:
            sim.sim_set_force_all_l2(0)
        lazy, _, loads_lazy, _ = run_sim(sim, orc, wl, None, slices, anchors, 2, 2, i % 2 == 0)
        assert (full == exp).all() and (lazy == exp).all()
        assert loads_lazy <= loads_full
    assert loads_lazy * 1.1 < loads_full          # sparse list: fewer early exits than on the 3 M list (there: 2.2x)


def test_sim_golden(sim, orc):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "bc_*.npz"))):
        g = np.load(f)
        wl, rank, slices, anchor = (np.ascontiguousarray(g[k]) for k in ("whitelist", "rank", "slices", "anchor"))
        got, _, _, _ = run_sim(sim, orc, wl, rank, slices, anchor, int(g["ed"]), 2, bool(g["three_prime"]))
        assert (got == g["result"]).all(), f


@pytest.mark.parametrize("umi_len", [12, 10, 14, 8])
def test_sim_umi(sim, orc, umi_len):
    umis, offs = workloads.umi_jobs(40 + umi_len, umi_len)
    exp, oo = orc.umi_matrix_batch(umis, offs, umi_len)
    got = np.zeros_like(exp)
    sim.sim_umi_dist(umis.ctypes.data, 16, umi_len, offs.ctypes.data, len(offs) - 1, got.ctypes.data, oo.ctypes.data)
    assert (got == exp).all()


def test_sim_umi_golden(sim):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "umi_*.npz"))):
        g = np.load(f)
        umis, offs, oo, exp = (np.ascontiguousarray(g[k]) for k in ("umis", "job_offsets", "out_offsets", "matrix"))
        got = np.zeros_like(exp)
        sim.sim_umi_dist(umis.ctypes.data, umis.shape[1], int(g["umi_len"]), offs.ctypes.data, len(offs) - 1, got.ctypes.data,
                         oo.ctypes.data)
        assert (got == exp).all()
