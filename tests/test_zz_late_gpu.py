"""GPU twins that were written after this round's GPU budget had run out: each has passed through its CPU twin (same chain, records from the
oracle or from the CPU replay of the kernel code instead of the kernel's), none has met the device yet.  They live in the module pytest
collects LAST so that a surprise here cannot cut short the `-x` run of the suites that have run on hardware."""
import numpy as np
import pytest

import __graft_entry__ as g
import test_ref_vectors as RV
from test_needleman import _bc_diff_check
from test_pipeline import GpuBackend, OracleBackend, check_sanity, run_pipeline
from test_usedlist import cases, same_collisions


pytestmark = pytest.mark.timeout(240, method="thread")      # never seen by a device: a wedged kernel must end the run, not sit in it


@pytest.fixture(scope="module")
def pk():
    p = g.load_package()
    p.build()
    return p


@pytest.mark.gpu
def test_gpu_mismatch_diff_matches_test_barcodes_bytecode(pk, ctx):
    def res_of(gene, allk, empk, sl, anchor, ed, pm, bail, slen):
        sets = pk.GuidedSets(ctx, gene, np.array([0, len(gene)], dtype=np.int64), 16, bc_flavour=True, all_keys=allk, all_ed=3, empty_keys=empk, empty_ed=2)
        return sets.match(sl, anchor, np.array([0], dtype=np.int32), ed, pm, 10, bailout=None if bail < 0 else bail, slice_len=slen)[0]
    _bc_diff_check(pk, res_of)


@pytest.mark.gpu
def test_gpu_pass1_to_pass2_list(pk, ctx):
    """the whole hand-over with the collision records of the GPU kernel: identical to the list the reference's class files kept"""
    for c, z, a, b, col in cases(pk):
        bc, cnt = z["barcodes"][a:b], z["counts"][a:b]
        table = pk.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, bc)
        got = pk.BarcodeDatasetColissionTester(ctx, table, int(z["ed"][c])).colissionsFromScan(bc)
        assert same_collisions(got, col), c                                                # slr_bc_collide == the reference's own doJob Matches
        keep, rank, _ = pk.used_merge_collisions(bc, cnt, got, int(z["min_count_fold"][c]), int(z["ed"][c]), int(z["cells_fold"][c]))
        assert np.array_equal(keep, z["kept"][a:b].astype(bool)), c


@pytest.mark.gpu
def test_gpu_pipeline_equals_oracle_pipeline(pkg, ctx, orc):
    a = run_pipeline(pkg, OracleBackend(orc))
    b = run_pipeline(pkg, GpuBackend(pkg, ctx))
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["lst"], b["lst"])
    assert np.array_equal(a["keep"], b["keep"]) and np.array_equal(a["rank"], b["rank"]) and np.array_equal(a["final"], b["final"])
    assert (a["rec1"] == b["rec1"]).all() and (a["rec2"] == b["rec2"]).all()                 # field by field (the records carry 4 bytes of struct padding)
    check_sanity(b, 300)


@pytest.mark.gpu
def test_gpu_umi_distance_wide_set(pkg, ctx):
    RV.gpu_umi_distance(pkg, ctx, RV.UMI_PAIR_FILES[1])


@pytest.mark.gpu
def test_gpu_cluster_one_hierarchical_wide_set(pkg, ctx):
    stats = RV.gpu_cluster_one_hierarchical(pkg, ctx, RV.HIER_FILES[1])
    assert stats[2] >= 400 and stats[0] > 3000, stats


@pytest.mark.gpu
def test_gpu_cluster_one_myclustering_wide_set(pkg, ctx):
    stats = RV.gpu_cluster_one_myclustering(pkg, ctx, RV.MYCLUST_FILES[1])
    assert stats[2] >= 90 and stats[0] > 9000, stats


@pytest.mark.gpu
def test_gpu_cluster_local_wide_set(pkg, ctx):
    RV.gpu_cluster_local(pkg, ctx, RV.CLUSTER_LOCAL_FILES[1])


@pytest.mark.gpu
def test_gpu_exact_lookup_wide_set(pkg, ctx):
    RV.gpu_exact_lookup(pkg, ctx, RV.EXACT_FILES[1])
