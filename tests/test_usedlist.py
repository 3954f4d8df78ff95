"""The hand-over between scanfastq's two passes (finalizeData): count filter, collision merge, ranks — the library's host code
(csrc/slr_usedlist.cpp) and the Python restatement oracle/pyref_usedlist.py against what the reference's own class files computed
(oracle/make_ref_usedlist.py -> tests/golden/ref_usedlist.npz: filterLowCounts and generateColissionMergedBCmap run by the interpreter on
Matches that the reference's BarcodeMatchTester.doJob produced).  Host arithmetic; the GPU twin takes the collision records from slr_bc_collide."""
import os

import numpy as np
import pytest

import __graft_entry__ as g
from oracle import pyref_usedlist as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_usedlist.npz")


@pytest.fixture(scope="module")
def pk():
    p = g.load_package()
    p.build()
    return p


def cases(pk):
    z = np.load(GOLD)
    off = z["offsets"]
    col = z["collide"].view(pk.COLLIDE_RESULT).reshape(-1)
    for c in range(len(off) - 1):
        a, b = int(off[c]), int(off[c + 1])
        yield c, z, a, b, col[a:b]


def test_merge_and_filter_match_reference_bytecode(pk):
    n_cases = n_chain = 0
    for c, z, a, b, col in cases(pk):
        bc, cnt = z["barcodes"][a:b], z["counts"][a:b]
        want = z["kept"][a:b].astype(bool)
        args = (int(z["min_count_fold"][c]), int(z["ed"][c]), int(z["cells_fold"][c]))
        keep, rank, flags = pk.used_merge_collisions(bc, cnt, col, *args)                  # slr_bc_used_merge_collisions
        assert np.array_equal(keep, want), c
        pkeep, prank, pflags = P.merge_collisions(bc, cnt, col, *args)                     # the independent restatement
        assert np.array_equal(pkeep, want) and np.array_equal(prank, rank) and pflags == flags, c
        # ranks: 1 = most reads, dense, only on kept barcodes
        assert (rank[~keep] == 0).all() and sorted(rank[keep].tolist()) == list(range(1, int(keep.sum()) + 1))
        o = np.argsort(rank[keep])
        assert (np.diff(cnt[keep][o]) <= 0).all()
        assert bool(flags & pk.UL_RANK_TIES) == (len(set(cnt[keep].tolist())) != int(keep.sum()))
        assert not flags & pk.UL_ORDER_UNPIN
        fk = pk.used_filter_low_counts(cnt, int(z["record_count"][c]))                     # slr_bc_used_filter_low_counts
        assert np.array_equal(fk, z["count_filter_keep"][a:b].astype(bool)), c
        assert np.array_equal(P.filter_low_counts(cnt, int(z["record_count"][c])), fk)
        # was the visiting order decisive?  (a barcode removed by a bigger one would have removed a third)
        naive = np.ones(len(bc), dtype=bool)
        idx = {int(x): i for i, x in enumerate(bc)}
        for i in range(len(bc)):
            for e in range(2):
                if col[i]["valid"] >> e & 1 and e + 1 <= args[1] and cnt[idx[int(col[i]["bc"][e])]] < cnt[i] // args[0]:
                    naive[idx[int(col[i]["bc"][e])]] = False
        n_chain += int((naive != (keep | (cnt < cnt[keep].max() // args[2]))).any())
        n_cases += 1
    assert n_cases >= 36 and n_chain >= 3          # the vectors include lists where "everybody removes" differs from the reference's order-dependent result


def test_vectors_pin_the_order_dependent_behaviour(pk):
    """switching one behaviour of the reference off changes the kept list of at least one recorded case"""
    def differing(**kw):
        bad = 0
        for c, z, a, b, col in cases(pk):
            keep, _, _ = P.merge_collisions(z["barcodes"][a:b], z["counts"][a:b], col, int(z["min_count_fold"][c]), int(z["ed"][c]), int(z["cells_fold"][c]), **kw)
            bad += not np.array_equal(keep, z["kept"][a:b].astype(bool))
        return bad
    assert differing() == 0
    assert differing(lazy=False) > 0, "a removed barcode removes nobody"
    assert differing(visiting="insertion") > 0, "the JDK HashMap's bin order, not the count order, decides chains"
    assert differing(strict=False) > 0, "count(c) < count(B) / minCountFold is strict"


def test_hashmap_order_model_agrees_with_interpreter_model(pk):
    """the bin layout used natively (and in pyref) against minijvm's JdkHashSet, the model the vectors were produced with"""
    from oracle import minijvm as J, make_ref_hier as H
    vm = H.HVM(H.JARS) if os.path.exists(H.JARS[0]) else None
    if vm is None:
        pytest.skip("reference jars not mounted (the GPU box): the vectors themselves pin the order")
    rng = np.random.default_rng(4)
    for _ in range(60):
        keys = list(dict.fromkeys(int(x) for x in rng.integers(0, 1 << 32, int(rng.integers(1, 300)))))
        hs = J.JdkHashSet(vm)
        for k in keys:
            hs.add(J.L(k))
        assert [int(x) for x in hs.items()] == P.hashmap_key_order(keys)[0]


def test_merge_refuses_bad_input(pk):
    bc = np.array([1, 2, 2], dtype=np.uint64)
    col = np.zeros(3, dtype=pk.COLLIDE_RESULT)
    with pytest.raises(pk.SiceloreGpuError) as e:
        pk.used_merge_collisions(bc, np.array([5, 4, 3]), col)
    assert e.value.code == pk.SLR_E_INVALID
    col = np.zeros(2, dtype=pk.COLLIDE_RESULT)
    col["valid"][0], col["bc"][0][0] = 1, 99                                               # names a barcode outside the list
    with pytest.raises(pk.SiceloreGpuError):
        pk.used_merge_collisions(np.array([1, 2], dtype=np.uint64), np.array([5, 4]), col)
    with pytest.raises(pk.SiceloreGpuError) as e:
        pk.used_merge_collisions(np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=pk.COLLIDE_RESULT))
    assert e.value.code == pk.SLR_E_REFERENCE_THROWS


def same_collisions(got, want):
    """the entries that exist agree (barcode and counters per ED level); absent entries are not compared"""
    if not np.array_equal(got["valid"] & 3, want["valid"] & 3):
        return False
    for e in range(2):
        m = (want["valid"] >> e & 1).astype(bool)
        for k in ("bc", "n_sub", "n_ins", "n_del"):
            if not np.array_equal(got[k][m, e], want[k][m, e]):
                return False
    return True


def test_oracle_collision_records_match_reference_dojob(pk, orc):
    """the collision records the merge consumes: C oracle (= what the GPU kernel is tested against) vs the Matches of the reference's own doJob"""
    for c, z, a, b, col in cases(pk):
        bc = z["barcodes"][a:b]
        got, _ = orc.collide_batch(orc.BarcodeSet(bc), bc, int(z["ed"][c]))
        assert same_collisions(got, col), c
        keep, _, _ = pk.used_merge_collisions(bc, z["counts"][a:b], got.view(pk.COLLIDE_RESULT), int(z["min_count_fold"][c]), int(z["ed"][c]), int(z["cells_fold"][c]))
        assert np.array_equal(keep, z["kept"][a:b].astype(bool)), c


def test_sim_collision_records_match_reference_dojob(pk, sim, orc):
    """the same lists (5 - 140 barcodes: tables far smaller than any other test builds) through the library's table builder and the CPU replay of
    the collision kernel's per-lane code — what the GPU twin below runs on the device"""
    import ctypes as C
    for c, z, a, b, col in cases(pk):
        bc = np.ascontiguousarray(z["barcodes"][a:b], dtype=np.uint64)
        got = np.zeros(len(bc), dtype=orc.COLLIDE_RESULT)
        loads = C.c_longlong(0)
        sim.sim_bc_collide(bc.ctypes.data, len(bc), 0, int(z["ed"][c]), bc.ctypes.data, len(bc), got.ctypes.data, C.byref(loads))
        assert same_collisions(got, col), c
