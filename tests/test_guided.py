"""Illumina-guided search: BCUMIEDtesterBase.matchesSeqEditDistance + the UMI / BC checkMatchWithTestSets + the
sorted().distinct() reduction of the match list.  CPU tests: the C oracle against the independent Python restatement, the
kernel's per-lane code (host simulation) against the oracle, frozen golden vectors, getmaxED.  GPU tests: the CUDA kernel
through the C ABI against the oracle."""
import os

import numpy as np
import pytest

import workloads
from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RAW_CAP = 48


def oracle_run(orc, w, L, ed, pm, post_len, bailout, bc, raw_cap=RAW_CAP):
    return orc.guided_batch(w["group_keys"], w["group_offsets"], w["slices"], w["anchor"], w["group_id"], ed, L, pm, post_len,
                            bailout=-1 if bailout is None else bailout, bc_flavour=bc, all_keys=w["all_keys"], all_ed=3,
                            empty_keys=w["empty_keys"], empty_ed=2, slice_len=w["slice_len"], raw_cap=raw_cap)


def sim_run(sim, orc, w, L, ed, pm, post_len, bailout, bc, raw_cap=RAW_CAP):
    n = len(w["slices"])
    out = np.zeros(n, dtype=orc.GUIDED_RESULT)
    raw = np.zeros((n, raw_cap), dtype=orc.GUIDED_HIT)
    edv = np.ascontiguousarray(np.broadcast_to(np.asarray(ed, dtype=np.int32), (n,)))
    ak, ek = w["all_keys"], w["empty_keys"]
    sim.sim_guided_batch(w["group_keys"].ctypes.data, w["group_offsets"].ctypes.data, len(w["group_offsets"]) - 1,
                         None if ak is None else ak.ctypes.data, 0 if ak is None else len(ak), 3,
                         None if ek is None else ek.ctypes.data, 0 if ek is None else len(ek), 2, int(bc), L, pm,
                         -1 if bailout is None else bailout, post_len, w["slices"].ctypes.data, 32, w["slice_len"], w["anchor"].ctypes.data,
                         w["group_id"].ctypes.data, edv.ctypes.data, n, out.ctypes.data, raw.ctypes.data, raw_cap)
    return out, raw


def assert_same(got, graw, exp, eraw, what):
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, "%s: %d records differ, first %d: got %s expected %s" % (what, len(bad), bad[0], got[bad[0]], exp[bad[0]])
    if graw is not None:
        ok = exp["flags"] == 0                           # the raw records of a read that throws are unspecified
        bad = np.nonzero((graw != eraw).any(axis=1) & ok)[0]
        assert len(bad) == 0, "%s: raw lists differ for %d reads, first %d" % (what, len(bad), bad[0])


CASES = [  # L, ed, pm, post_len, bailout, bc_flavour, n queries (python restatement / simulation)
    (12, 0, 2, 4, None, False, 60, 300),
    (12, 1, 2, 5, None, False, 60, 300),
    (12, 2, 2, 6, None, False, 30, 200),
    (12, 2, 1, 5, 1, False, 30, 200),
    (16, 1, 2, 10, None, True, 40, 300),
    (16, 2, 2, 10, 2, True, 20, 150),
    (16, 2, 1, 10, 1, True, 20, 150),
    (10, 3, 1, 6, None, False, 3, 12),
    (16, 3, 0, 10, 2, True, 2, 8),
    (8, 4, 0, 6, None, False, 1, 2),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "L%d_ed%d_pm%d_bail%s_%s" % (c[0], c[1], c[2], c[4], "bc" if c[5] else "umi"))
def test_oracle_vs_python_restatement(orc, case):
    """every field of the result, the raw list (order, counters, flags) and the exception cases agree between slr_oracle.c and pyref.py"""
    L, ed, pm, post_len, bailout, bc, nq, _ = case
    w = workloads.guided(100 + ed, L, nq, ed, pm, post_len, bc, skew=(ed == 2))
    res, raw, probes = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
    n_exc = n_hit = 0
    for q in range(nq):
        gid = int(w["group_id"][q])
        grp = set(w["groups"][gid]) if 0 <= gid < len(w["groups"]) else set()
        sets = dict(umis=grp) if not bc else dict(gene=grp or None, all_bcs=set(int(x) for x in w["all_keys"]), all_ed=3,
                                                  empty=set(int(x) for x in w["empty_keys"]), empty_ed=2)
        r = res[q]
        try:
            praw, plst = pyref.guided_query(bytes(w["slices"][q][:w["slice_len"]]), int(w["anchor"][q]), L, ed, pm, post_len,
                                            bailout=bailout, bc_flavour=bc, **sets)
        except pyref.JavaException:
            assert r["flags"] & orc.G_EXCEPTION and r["n_raw"] == 0 and r["n_distinct"] == 0
            n_exc += 1
            continue
        assert not r["flags"]
        assert r["n_raw"] == len(praw)
        for i, h in enumerate(praw[:RAW_CAP]):
            x = raw[q, i]
            assert (int(x["seq"]), x["n_sub"], x["n_ins"], x["n_del"], x["offset"], x["where"], x["level"]) == \
                   (h["seq"], h["n_sub"], h["n_ins"], h["n_del"], h["offset"], h["where"], h["level"])
        assert r["n_distinct"] == min(2, len(plst))
        for i, h in enumerate(plst[:2]):
            assert (int(r["seq"][i]), r["n_sub"][i], r["n_ins"][i], r["n_del"][i], r["offset"][i], r["where"][i]) == \
                   (h["seq"], h["n_sub"], h["n_ins"], h["n_del"], h["offset"], h["where"])
        ge = [h["n_sub"] + h["n_ins"] + h["n_del"] for h in praw if h["where"] & 1]
        assert r["min_err_gene"] == (min(ge) if ge else 2147483647)
        n_hit += len(praw) > 0
    assert n_hit > 0 or nq < 3


@pytest.mark.parametrize("case", CASES, ids=lambda c: "L%d_ed%d_pm%d_bail%s_%s" % (c[0], c[1], c[2], c[4], "bc" if c[5] else "umi"))
def test_sim_vs_oracle(sim, orc, case):
    L, ed, pm, post_len, bailout, bc, _, nq = case
    for skew in (False, True):
        w = workloads.guided(200 + 7 * ed + skew, L, nq, ed, pm, post_len, bc, skew=skew)
        exp, eraw, _ = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
        got, graw = sim_run(sim, orc, w, L, ed, pm, post_len, bailout, bc)
        assert_same(got, graw, exp, eraw, "host simulation")
        assert (exp["n_raw"] > 0).sum() > 0 or nq < 10


@pytest.mark.parametrize("L,post_len,bc", [(12, 6, False), (16, 10, True)])
def test_sim_ed4_full_length(sim, orc, L, post_len, bc):
    """the deepest supported search (IlluminaUMIanalyzer.java:L75: ed < 5) at real UMI / barcode lengths: ~0.5-2 M probes per window"""
    w = workloads.guided(900 + L, L, 6, 4, 1, post_len, bc)
    exp, eraw, _ = oracle_run(orc, w, L, 4, 1, post_len, 2 if bc else None, bc)
    got, graw = sim_run(sim, orc, w, L, 4, 1, post_len, 2 if bc else None, bc)
    assert_same(got, graw, exp, eraw, "ED 4")
    assert (exp["n_raw"] > 0).any()


def test_candidate_filter_is_conservative(sim, orc):
    """the last-level batch of a node is skipped when no candidate of a small group can be one of its children: every child the engine
    can create passes the filter (exhaustive over positions x operations of random nodes), skipping changes no field of any record,
    and it does skip most batches"""
    for L, post_len in ((12, 6), (16, 10), (8, 5), (5, 4)):
        assert sim.sim_guided_filter_violations(L, post_len, 20000, L) == 0
    w = workloads.guided(55, 12, 200, 2, 2, 6, False, skew=True)
    sim.sim_guided_filter_skips(1)
    got, graw = sim_run(sim, orc, w, 12, 2, 2, 6, None, False)
    skips = sim.sim_guided_filter_skips(1)
    sim.sim_guided_set_filter(0)
    try:
        ref, rraw = sim_run(sim, orc, w, 12, 2, 2, 6, None, False)
    finally:
        sim.sim_guided_set_filter(1)
    assert sim.sim_guided_filter_skips(1) == 0
    assert_same(got, graw, ref, rraw, "filter on vs off")
    assert skips > 10000
    # far-node batches one level above the last (ED 3: the level-2 nodes; with a bailout; BC flavour at ED 4 where the lists are out of reach)
    for L, ed, pm, post_len, bailout, bc, nq in ((12, 3, 1, 6, None, False, 24), (12, 3, 1, 6, 2, False, 24), (10, 4, 0, 6, None, False, 6),
                                                 (16, 4, 0, 10, 2, True, 4), (12, 2, 2, 6, None, False, 150)):
        w = workloads.guided(60 + ed, L, nq, ed, pm, post_len, bc, skew=True, group_sizes=(0, 1, 2, 5, 12, 25))
        sim.sim_guided_far_nodes(1)
        got, graw = sim_run(sim, orc, w, L, ed, pm, post_len, bailout, bc)
        far = sim.sim_guided_far_nodes(1)
        sim.sim_guided_set_filter(0)
        try:
            ref, rraw = sim_run(sim, orc, w, L, ed, pm, post_len, bailout, bc)
        finally:
            sim.sim_guided_set_filter(1)
        assert_same(got, graw, ref, rraw, "far-node batches on vs off")
        exp, eraw, _ = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
        assert_same(got, graw, exp, eraw, "far-node batches vs oracle")
        assert far > 0 or bc, (L, ed, bc)      # BC flavour: the all-passed list (maxEDtoCheckBCAll10xBCs 3) is still probed at level 3


def sweep_cases(n=36):
    """seeded sweep over the whole parameter space the ABI accepts: sequence length 4..16, +-0..3, ED 0..3, bailout null / 0..3, post length,
    both flavours, group sizes around the filter's 64-slot limit"""
    rng = np.random.default_rng(2026)
    for it in range(n):
        L = int(rng.integers(4, 17))
        ed = int(rng.integers(0, 4 if L <= 12 else 3))
        pm = int(rng.integers(0, 4))
        post_len = int(rng.integers(ed + 1, ed + 6))
        while 2 * pm + L + post_len + 2 > 32:
            pm -= 1
        bc = bool(it % 3 == 0)
        bailout = None if it % 4 == 0 else int(rng.integers(0, 4))
        sizes = (0, 1, int(rng.integers(2, 9)), int(rng.integers(9, 31)), 33, int(rng.integers(34, 90)))
        w = workloads.guided(5000 + it, L, 14 if ed >= 3 else 40, ed, pm, post_len, bc, skew=bool(it % 2), group_sizes=sizes,
                             n_all=int(rng.integers(1, 40)), n_empty=int(rng.integers(1, 80)))
        yield it, w, L, ed, pm, post_len, bailout, bc


def test_sim_random_parameter_sweep(sim, orc):
    hits = 0
    for it, w, L, ed, pm, post_len, bailout, bc in sweep_cases():
        exp, eraw, _ = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
        got, graw = sim_run(sim, orc, w, L, ed, pm, post_len, bailout, bc)
        assert_same(got, graw, exp, eraw, "sweep %d: L %d ed %d pm %d post %d bail %s bc %s" % (it, L, ed, pm, post_len, bailout, bc))
        hits += int((exp["n_raw"] > 0).sum())
    assert hits > 300


def test_sim_mixed_edit_distances(sim, orc):
    """dynamic ED: every read carries its own maxEDdyn; the stamped visited table is shared by windows of different table sizes"""
    w = workloads.guided(77, 12, 120, 2, 2, 6, False)
    rng = np.random.default_rng(3)
    ed = rng.integers(0, 3, 120).astype(np.int32)
    ed[:4] = 3
    exp, eraw, _ = oracle_run(orc, w, 12, ed, 2, 6, None, False)
    got, graw = sim_run(sim, orc, w, 12, ed, 2, 6, None, False)
    assert_same(got, graw, exp, eraw, "mixed ED")


def test_duplicates_and_order_properties(orc):
    """list semantics the consumers rely on: entries of one window are in traversal order (offset blocks 0,-1,+1,…), a candidate hit at
    the last level is reported once per path (duplicates), the reduced list is sorted by (errors, |offset|) and distinct"""
    L, ed, pm, post_len = 12, 2, 1, 5
    w = workloads.guided(5, L, 150, 2, pm, post_len, False, special=False)
    res, raw, _ = oracle_run(orc, w, L, ed, pm, post_len, None, False, raw_cap=256)
    order = {0: 0, -1: 1, 1: 2}
    dup = 0
    for q in range(150):
        r = res[q]
        if r["flags"] or r["n_raw"] == 0:
            continue
        hits = raw[q][:min(int(r["n_raw"]), 256)]
        blocks = [order[int(o)] for o in hits["offset"]]
        assert blocks == sorted(blocks)
        dup += len(hits) - len({(int(h["seq"]), int(h["offset"])) for h in hits})
        errs = (hits["n_sub"] + hits["n_ins"] + hits["n_del"]).astype(int)
        key = errs * 8 + np.abs(hits["offset"].astype(int))
        if int(r["n_raw"]) <= 256:
            assert int(r["n_sub"][0]) + int(r["n_ins"][0]) + int(r["n_del"][0]) == errs.min()
            first = int(np.argmin(key))                  # stable: first entry holding the smallest key
            assert int(r["seq"][0]) == int(hits["seq"][first]) and int(r["offset"][0]) == int(hits["offset"][first])
            if r["n_distinct"] == 2:
                assert int(r["seq"][1]) != int(r["seq"][0])
            else:
                assert len(set(int(s) for s in hits["seq"])) == 1
    assert dup > 0


def test_bc_flag_inheritance(orc):
    """BCnucTwoBitPerBaseEDtester puts the GENE bit on the NODE (java:L76) and the copy constructor hands it to every descendant:
    an all-passed hit below a gene hit carries GENE | ALL and sorts with score 3"""
    L, post_len = 16, 10
    gene = workloads.g_pack("AGCTAGCTAGCTAGCT")
    other = workloads.g_pack("AGCTAGCTAGCTAGCA")       # one substitution (last base) away from `gene`
    far = workloads.g_pack("TTTTAGCTAGCTAGCT")
    s = ("AGCTAGCTAGCTAGCT" + "GGGGGGGGGGGGGGGG")[:32]
    sl = np.frombuffer(s.encode(), dtype=np.uint8).reshape(1, 32).copy()
    res, raw, _ = orc.guided_batch(np.array([gene], dtype=np.uint64), np.array([0, 1], dtype=np.int64), sl, np.array([0], dtype=np.int32),
                                   np.array([0], dtype=np.int32), 1, L, 0, post_len, bc_flavour=True,
                                   all_keys=np.array([other, far], dtype=np.uint64), all_ed=3, empty_keys=None, slice_len=32, raw_cap=8)
    assert res[0]["n_raw"] == 2
    assert raw[0, 0]["where"] == orc.W_GENE and int(raw[0, 0]["seq"]) == gene
    assert raw[0, 1]["where"] == (orc.W_GENE | orc.W_ALL) and int(raw[0, 1]["seq"]) == other      # inherited from the root
    # without the gene hit the same all-passed entry carries ALL only
    res2, raw2, _ = orc.guided_batch(np.array([far], dtype=np.uint64), np.array([0, 1], dtype=np.int64), sl, np.array([0], dtype=np.int32),
                                     np.array([0], dtype=np.int32), 1, L, 0, post_len, bc_flavour=True,
                                     all_keys=np.array([other], dtype=np.uint64), all_ed=3, empty_keys=None, slice_len=32, raw_cap=8)
    assert res2[0]["n_raw"] == 1 and raw2[0, 0]["where"] == orc.W_ALL


def test_root_level_is_one(orc):
    """the root is probed with currentlevel 1 (BCUMIEDtesterBase.java:L82): with maxEDtoCheckBCAll10xBCs = 0 an exact all-passed
    match is NOT reported"""
    L, post_len = 16, 10
    k = workloads.g_pack("AGCTAGCTAGCTAGCT")
    sl = np.frombuffer(("AGCTAGCTAGCTAGCT" + "G" * 16).encode(), dtype=np.uint8).reshape(1, 32).copy()
    args = (np.zeros(0, dtype=np.uint64), np.array([0, 0], dtype=np.int64), sl, np.array([0], dtype=np.int32), np.array([0], dtype=np.int32))
    r0, _, _ = orc.guided_batch(*args, 0, L, 0, post_len, bc_flavour=True, all_keys=np.array([k], dtype=np.uint64), all_ed=0, slice_len=32)
    r1, _, _ = orc.guided_batch(*args, 0, L, 0, post_len, bc_flavour=True, all_keys=np.array([k], dtype=np.uint64), all_ed=1, slice_len=32)
    assert r0[0]["n_raw"] == 0 and r1[0]["n_raw"] == 1


def test_deletion_at_last_position(orc):
    """unlike BarcodeMatchTester.doJob, deletions also run at the last position (BCUMIEDtesterBase.java:L113-L119): the last base is
    replaced by the first post base and counted as an insertion"""
    L = 12
    cand = workloads.g_pack("AGCTAGCTAGCG")             # window with its last base replaced by the post base G
    sl = np.frombuffer(("AGCTAGCTAGCT" + "GAAAAAAAAAAAAAAAAAAA").encode(), dtype=np.uint8).reshape(1, 32).copy()
    res, raw, _ = orc.guided_batch(np.array([cand], dtype=np.uint64), np.array([0, 1], dtype=np.int64), sl, np.array([0], dtype=np.int32),
                                   np.array([0], dtype=np.int32), 1, L, 0, 4, slice_len=32, raw_cap=8)
    kinds = {(int(h["n_sub"]), int(h["n_ins"]), int(h["n_del"])) for h in raw[0][:int(res[0]["n_raw"])]}
    assert (0, 1, 0) in kinds and (1, 0, 0) in kinds     # once as a substitution, once as the last-position deletion


def test_golden_guided(sim, orc):
    """frozen known-answer vectors (tests/golden/guided_*.npz, written by tests/golden/make_golden.py): oracle and simulation"""
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "guided_*.npz")))
    assert files, "golden guided vectors missing"
    for f in files:
        z = np.load(f)
        w = {k: (z[k] if k in z.files and z[k].size else None) for k in ("all_keys", "empty_keys")}
        w.update(group_keys=z["group_keys"], group_offsets=z["group_offsets"], slices=z["slices"], anchor=z["anchor"], group_id=z["group_id"],
                 slice_len=int(z["slice_len"]))
        L, pm, post_len, bailout, bc = (int(z[k]) for k in ("L", "pm", "post_len", "bailout", "bc"))
        bailout = None if bailout < 0 else bailout
        exp = z["result"].view(orc.GUIDED_RESULT).reshape(-1)
        eraw = z["raw"].view(orc.GUIDED_HIT).reshape(len(exp), -1)
        got, graw, _ = oracle_run(orc, w, L, z["ed"], pm, post_len, bailout, bc, raw_cap=eraw.shape[1])
        assert_same(got, graw, exp, eraw, os.path.basename(f) + " oracle")
        got, graw = sim_run(sim, orc, w, L, z["ed"], pm, post_len, bailout, bc, raw_cap=eraw.shape[1])
        assert_same(got, graw, exp, eraw, os.path.basename(f) + " simulation")


# columns of Jar/umiMaxEditDistances.xml (length 12, 1 %) and Jar/bcMaxEditDistances.xml (length 16, 6 %)
UMI12_1PCT = [16777216, 2849, 103, 8, 2]
BC16_6PCT = [4294967296, 600000, 158172, 7031, 474]


def java_getmaxed(col, count, pm, cap):
    """DynamicEditDistances.getmaxED (java:L93-L98) literally: filter(maxBarcodes >= count*(2pm+1)), max by key, cap"""
    ok = [e for e, v in enumerate(col) if v >= count * (2 * pm + 1)]
    if not ok:
        return -1                                        # Optional.get() throws NoSuchElementException
    r = max(ok)
    return min(r, cap) if cap is not None and r > cap else r


def test_dyn_max_ed(pkg):
    lib = pkg.gpu_lib()
    for col in (UMI12_1PCT, BC16_6PCT):
        arr = np.array(col, dtype=np.int64)
        for count in (1, 2, 3, 20, 21, 569, 570, 100000, 5000000):
            for pm in (0, 1, 2):
                for cap in (None, 0, 2, 4):
                    got = lib.slr_dyn_max_ed(arr.ctypes.data, len(arr), count, pm, -1 if cap is None else cap)
                    assert got == java_getmaxed(col, count, pm, cap), (col, count, pm, cap)
    assert lib.slr_dyn_max_ed(np.array([1], dtype=np.int64).ctypes.data, 1, 5, 2, -1) == -1      # Optional.get() would throw


# ------------------------------------------------------------------------------------------------------------- GPU
def gpu_run(pkg, ctx, w, L, ed, pm, post_len, bailout, bc, raw_cap=RAW_CAP):
    sets = pkg.GuidedSets(ctx, w["group_keys"], w["group_offsets"], L, bc_flavour=bc, all_keys=w["all_keys"], all_ed=3,
                          empty_keys=w["empty_keys"], empty_ed=2)
    return sets.match(w["slices"], w["anchor"], w["group_id"], ed, pm, post_len, bailout=bailout, slice_len=w["slice_len"], raw_cap=raw_cap)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "L%d_ed%d_pm%d_bail%s_%s" % (c[0], c[1], c[2], c[4], "bc" if c[5] else "umi"))
def test_gpu_vs_oracle(pkg, ctx, orc, case):
    L, ed, pm, post_len, bailout, bc, _, nq = case
    nq = nq * (20 if ed <= 2 else 4)
    for skew in (False, True):
        w = workloads.guided(300 + 7 * ed + skew, L, nq, ed, pm, post_len, bc, skew=skew)
        exp, eraw, _ = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
        got, graw = gpu_run(pkg, ctx, w, L, ed, pm, post_len, bailout, bc)
        assert_same(got, graw, exp, eraw, "GPU")


@pytest.mark.gpu
@pytest.mark.parametrize("L,post_len,bc", [(12, 6, False), (16, 10, True)])
def test_gpu_ed4_full_length(pkg, ctx, orc, L, post_len, bc):
    w = workloads.guided(900 + L, L, 64, 4, 1, post_len, bc)
    exp, eraw, _ = oracle_run(orc, w, L, 4, 1, post_len, 2 if bc else None, bc)
    got, graw = gpu_run(pkg, ctx, w, L, 4, 1, post_len, 2 if bc else None, bc)
    assert_same(got, graw, exp, eraw, "GPU ED 4")


@pytest.mark.gpu
def test_gpu_random_parameter_sweep(pkg, ctx, orc):
    for it, w, L, ed, pm, post_len, bailout, bc in sweep_cases():
        exp, eraw, _ = oracle_run(orc, w, L, ed, pm, post_len, bailout, bc)
        got, graw = gpu_run(pkg, ctx, w, L, ed, pm, post_len, bailout, bc)
        assert_same(got, graw, exp, eraw, "GPU sweep %d: L %d ed %d pm %d post %d bail %s bc %s" % (it, L, ed, pm, post_len, bailout, bc))


@pytest.mark.gpu
def test_gpu_mixed_edit_distances_and_no_raw(pkg, ctx, orc):
    w = workloads.guided(78, 12, 3000, 2, 2, 6, False)
    ed = np.random.default_rng(4).integers(0, 3, 3000).astype(np.int32)
    ed[:16] = 3
    exp, _, _ = oracle_run(orc, w, 12, ed, 2, 6, None, False)
    got, graw = gpu_run(pkg, ctx, w, 12, ed, 2, 6, None, False, raw_cap=0)
    assert graw is None
    assert_same(got, None, exp, None, "GPU mixed ED")


@pytest.mark.gpu
def test_gpu_golden_guided(pkg, ctx, orc):
    import glob
    for f in sorted(glob.glob(os.path.join(GOLDEN, "guided_*.npz"))):
        z = np.load(f)
        w = {k: (z[k] if k in z.files and z[k].size else None) for k in ("all_keys", "empty_keys")}
        w.update(group_keys=z["group_keys"], group_offsets=z["group_offsets"], slices=z["slices"], anchor=z["anchor"], group_id=z["group_id"],
                 slice_len=int(z["slice_len"]))
        L, pm, post_len, bailout, bc = (int(z[k]) for k in ("L", "pm", "post_len", "bailout", "bc"))
        exp = z["result"].view(orc.GUIDED_RESULT).reshape(-1)
        eraw = z["raw"].view(orc.GUIDED_HIT).reshape(len(exp), -1)
        got, graw = gpu_run(pkg, ctx, w, L, z["ed"], pm, post_len, None if bailout < 0 else bailout, bc, raw_cap=eraw.shape[1])
        assert_same(got, graw, exp, eraw, os.path.basename(f) + " GPU")


@pytest.mark.gpu
def test_gpu_guided_refusals(pkg, ctx):
    w = workloads.guided(1, 12, 10, 1, 1, 5, False)
    sets = pkg.GuidedSets(ctx, w["group_keys"], w["group_offsets"], 12)
    with pytest.raises(pkg.SiceloreGpuError):      # ed 5: "ED >5 will require huge CPU time" (IlluminaUMIanalyzer.java:L75)
        sets.match(w["slices"], w["anchor"], w["group_id"], 5, 1, 8)
    with pytest.raises(pkg.SiceloreGpuError):      # post sequence shorter than ed + 1
        sets.match(w["slices"], w["anchor"], w["group_id"], 3, 1, 3)
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.GuidedSets(ctx, w["group_keys"], w["group_offsets"], 17)
