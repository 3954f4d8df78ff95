"""Pass-1 barcode collision tester (seam S3): BarcodeDatasetColissionTester.submitSeq runs the barcode engine
with skipFullMatches = true, postSeq = null, doNextLevelIfMatchFound = false for every used barcode against the list itself
(BarcodeDatasetColissionTester.java:L212-L229).  Oracle vs the independent Python restatement, the kernel's per-lane code on
the CPU (tests/host_sim) vs the oracle, frozen golden vectors, and the CUDA kernel through the C ABI (-m gpu)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import workloads
from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pyref_collide(wl, q, ed):
    s = set(int(x) for x in wl)
    t = pyref.BarcodeMatchTester(int(q), 16, ed, True, True, s, 0, None, False)
    return {m.ed: m for m in (t.doJob() or [])}


@pytest.mark.parametrize("skew", [False, True])
@pytest.mark.parametrize("ed", [1, 2])
def test_oracle_vs_pyref(orc, ed, skew):
    wl = workloads.used_list(50 + ed + 10 * skew, 60, skew)
    exp, probes = orc.collide_batch(orc.BarcodeSet(wl), wl, ed)
    assert probes > 0
    seen = 0
    for i, q in enumerate(wl):
        ms = pyref_collide(wl, q, ed)
        for lv in (1, 2):
            has = bool((exp["valid"][i] >> (lv - 1)) & 1)
            assert has == (lv in ms), (i, lv)
            if has:
                m = ms[lv]
                assert int(exp["bc"][i][lv - 1]) == m.bc and int(exp["bc"][i][lv - 1]) != int(q)
                assert (exp["n_sub"][i][lv - 1], exp["n_ins"][i][lv - 1], exp["n_del"][i][lv - 1]) == (m.nSub, m.nIns, m.nDel)
                seen += 1
    assert seen > 20


def test_oracle_collision_semantics(orc):
    """skipFullMatches: a barcode never collides with itself; postSeq = null: a deletion may append any base; doNext = false:
    below a SUB child that missed nothing is searched, below an INS/DEL child that hit nothing is searched."""
    P = pyref.pack
    a = "AGCTTGACCATGGTCA"
    # alone in the list: no entry at all (ED 0 is skipped, nothing else there)
    r, _ = orc.collide_batch(orc.BarcodeSet(np.array([P(a)], dtype=np.uint64)), np.array([P(a)], dtype=np.uint64), 2)
    assert r["valid"][0] == 0
    # deletion at position 3 + appended G / T: both are ED-1 neighbours without a post sequence
    for tail in "AGCT":
        b = a[:3] + a[4:] + tail
        wl = np.array([P(a), P(b)], dtype=np.uint64)
        r, _ = orc.collide_batch(orc.BarcodeSet(wl), wl[:1], 1)
        assert r["valid"][0] == 1 and int(r["bc"][0][0]) == P(b) and r["n_ins"][0][0] == 1
    # two substitutions: the ED-2 neighbour is only reached through a SUB child that is itself in the list
    c1 = "T" + a[1:]
    c2 = "T" + a[1:8] + "G" + a[9:]
    wl = np.array([P(a), P(c2)], dtype=np.uint64)
    r, _ = orc.collide_batch(orc.BarcodeSet(wl), wl[:1], 2)
    assert r["valid"][0] == 0                                   # c1 missing: the substitution branch is never expanded (L268)
    wl = np.array([P(a), P(c1), P(c2)], dtype=np.uint64)
    r, _ = orc.collide_batch(orc.BarcodeSet(wl), wl[:1], 2)
    assert r["valid"][0] == 3 and int(r["bc"][0][0]) == P(c1) and int(r["bc"][0][1]) == P(c2) and r["n_sub"][0][1] == 2
    # insertion then substitution: reached through an INS child that is NOT in the list
    d1 = a[:5] + "T" + a[5:15]
    d2 = d1[:10] + ("A" if d1[10] != "A" else "C") + d1[11:]
    wl = np.array([P(a), P(d2)], dtype=np.uint64)
    r, _ = orc.collide_batch(orc.BarcodeSet(wl), wl[:1], 2)
    assert r["valid"][0] == 2 and int(r["bc"][0][1]) == P(d2)
    wl = np.array([P(a), P(d1), P(d2)], dtype=np.uint64)      # ... and cut off once that child hits (L295)
    r, _ = orc.collide_batch(orc.BarcodeSet(wl), wl[:1], 2)
    assert (r["valid"][0] & 1) and int(r["bc"][0][0]) == P(d1)
    assert not ((r["valid"][0] & 2) and int(r["bc"][0][1]) == P(d2))     # (the ED-2 slot goes to d1 again, reached below a missed INS)


def run_sim(sim, wl, queries, ed):
    from oracle import orc
    got = np.zeros(len(queries), dtype=orc.COLLIDE_RESULT)
    loads = C.c_longlong(0)
    sim.sim_bc_collide(wl.ctypes.data, len(wl), 0, ed, queries.ctypes.data, len(queries), got.ctypes.data, C.byref(loads))
    return got, loads.value


@pytest.mark.parametrize("skew", [False, True])
@pytest.mark.parametrize("ed", [0, 1, 2])
def test_sim_vs_oracle(sim, orc, ed, skew):
    wl = workloads.used_list(300 + ed + 10 * skew, 500, skew)
    q = np.concatenate([wl, np.random.default_rng(1).integers(0, 1 << 32, 200, dtype=np.uint64),
                        np.array([(1 << 40) | 5], dtype=np.uint64)])            # + queries outside the list, one with garbage high bits
    exp, probes = orc.collide_batch(orc.BarcodeSet(wl), q, ed)
    got, loads = run_sim(sim, wl, q, ed)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    if ed:
        assert (exp["valid"] & 1).sum() > 100
    if ed == 2:
        assert loads * 3 < probes                       # level 1: one exact lookup per mutant like the reference; level 2: 21 loads per node


def test_golden(sim, orc):
    fs = sorted(glob.glob(os.path.join(GOLDEN, "collide_*.npz")))
    assert fs
    for f in fs:
        g = np.load(f)
        wl = np.ascontiguousarray(g["whitelist"])
        exp, _ = orc.collide_batch(orc.BarcodeSet(wl), wl, int(g["ed"]))
        assert (exp == g["result"]).all(), f
        got, _ = run_sim(sim, wl, wl, int(g["ed"]))
        assert (got == g["result"]).all(), f


# ---- pass-1 exact lookup (UsedCellBCListGenerator$Worker, UsedCellBCListGenerator.java:L206-L232) -------------------
def exact_case(seed, three_prime):
    """adversarial slices whose offset-0 windows are partly planted in the list; anchors near both slice ends (no flank)"""
    _, slices, anchors, wl = workloads.adversarial(seed, three_prime, 400, nrand=100)
    rng = np.random.default_rng(seed)
    anchors = rng.integers(-1, 18, size=len(slices)).astype(np.int32)
    keys = set(int(x) for x in wl)
    for i in range(0, len(slices), 2):                          # plant every other read's own window
        a = int(anchors[i])
        if 0 <= a <= 16:
            s = bytes(slices[i, a:a + 16]).decode("latin1")
            if all(c in "ACGTacgt" for c in s):
                keys.add(pyref.pack(workloads.rcs(s.upper()) if three_prime else s.upper()))
    wl = np.array(sorted(keys), dtype=np.uint64)
    return slices, anchors, wl


@pytest.mark.parametrize("three_prime", [True, False])
def test_exact_lookup_sim_vs_oracle(sim, orc, three_prime):
    slices, anchors, wl = exact_case(11 + three_prime, three_prime)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    exp = orc.exact_lookup_batch(orc.BarcodeSet(wl, rank), slices, anchors, three_prime)
    assert (exp["flags"] & 1).sum() > 100 and (exp["flags"] & 2).sum() > 5
    sim.sim_set_need_post.argtypes = [C.c_int]
    got = np.zeros(len(slices), dtype=orc.BC_RESULT)
    counts = np.zeros(len(wl) * 3, dtype=np.uint64)
    loads = C.c_longlong(0)
    st = np.zeros(4, dtype=np.int64)
    try:
        sim.sim_set_need_post(0)
        sim.sim_bc_assign(wl.ctypes.data, rank.ctypes.data, len(wl), 0, 0, 0, int(three_prime), slices.ctypes.data, 32, 32, None,
                          anchors.ctypes.data, len(slices), got.ctypes.data, counts.ctypes.data, C.byref(loads), st.ctypes.data)
    finally:
        sim.sim_set_need_post(1)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    assert counts.reshape(-1, 3)[:, 0].sum() == (exp["flags"] & 1).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("three_prime", [True, False])
def test_exact_lookup_gpu(pkg, orc, ctx, three_prime):
    slices, anchors, wl = exact_case(21 + three_prime, three_prime)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    exp = orc.exact_lookup_batch(orc.BarcodeSet(wl, rank), slices, anchors, three_prime)
    table = pkg.BarcodesMapForBCfinding(ctx, wl, rank)
    gen = pkg.UsedCellBCListGenerator(ctx, table, three_prime)
    got = gen.addFastqs(slices, anchors)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    keys, cnt = gen.unfilteredUsedBarcodeMap()
    ok = (exp["flags"] & 1) == 1
    ek, ec = np.unique(exp["bc"][ok], return_counts=True)
    order = np.argsort(keys)
    assert (keys[order] == ek).all() and (cnt[order] == ec).all()
    # synthetic reads of configs[1]: the pass-1 list = barcodes seen without error at the predicted position
    wl = pkg.synth_whitelist(737280, 737)
    sl, an, truth = pkg.synth_reads(wl, 200000, seed=1, three_prime=three_prime)
    table = pkg.BarcodesMapForBCfinding(ctx, wl)
    got = pkg.UsedCellBCListGenerator(ctx, table, three_prime).addFastqs(sl, an)
    exp = orc.exact_lookup_batch(orc.BarcodeSet(wl), sl, an, three_prime)
    assert (got == exp).all()
    assert 0.2 < ((got["flags"] & 1) == 1).mean() < 0.6


# ---- the CUDA kernel through the C ABI ------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("skew", [False, True])
@pytest.mark.parametrize("ed", [0, 1, 2])
def test_gpu_vs_oracle(pkg, orc, ctx, ed, skew):
    wl = workloads.used_list(700 + ed + 10 * skew, 3000, skew)
    table = pkg.BarcodesMapForBCfinding(ctx, wl)
    got = pkg.BarcodeDatasetColissionTester(ctx, table, ed).colissionsFromScan()
    exp, _ = orc.collide_batch(orc.BarcodeSet(wl), wl, ed)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    if ed:
        assert (exp["valid"] & 1).sum() > 1000


@pytest.mark.gpu
def test_gpu_golden_edges_and_realistic_list(pkg, orc, ctx):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "collide_*.npz"))):
        g = np.load(f)
        wl = np.ascontiguousarray(g["whitelist"])
        table = pkg.BarcodesMapForBCfinding(ctx, wl)
        got = pkg.BarcodeDatasetColissionTester(ctx, table, int(g["ed"])).colissionsFromScan()
        assert (got == g["result"]).all(), f
    # empty query list, queries that are not in the list, mergeBCsED 3 refused
    wl = workloads.used_list(9, 200)
    table = pkg.BarcodesMapForBCfinding(ctx, wl)
    t = pkg.BarcodeDatasetColissionTester(ctx, table, 2)
    assert len(t.colissionsFromScan(wl[:0])) == 0
    q = np.random.default_rng(2).integers(0, 1 << 32, 500, dtype=np.uint64)
    exp, _ = orc.collide_batch(orc.BarcodeSet(wl), q, 2)
    assert (t.colissionsFromScan(q) == exp).all()
    with pytest.raises(pkg.SiceloreGpuError) as e:
        pkg.BarcodeDatasetColissionTester(ctx, table, 3).colissionsFromScan()
    assert e.value.code == pkg.SLR_E_UNSUPPORTED
    # a used list of the size pass 1 produces: 60 000 whitelist barcodes + their error children, vs the oracle
    base = pkg.synth_whitelist(60000, 11)
    rng = np.random.default_rng(5)
    kids = base[rng.integers(0, len(base), 40000)] ^ (np.uint64(1) << rng.integers(0, 32, 40000).astype(np.uint64))
    wl = np.unique(np.concatenate([base, kids]))
    table = pkg.BarcodesMapForBCfinding(ctx, wl)
    got = pkg.BarcodeDatasetColissionTester(ctx, table, 2).colissionsFromScan()
    exp, _ = orc.collide_batch(orc.BarcodeSet(wl), wl, 2)
    assert (got == exp).all()
    assert (got["valid"] & 1).sum() > 30000
