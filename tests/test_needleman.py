"""The alignment step that decides MORE_THAN_ONE_MATCH in the Illumina-guided search (IlluminaBarcodeUMIAnalyzerBase.getBestAndSecondBCorUMI,
…java:L66-L86): the library's host code (csrc/slr_needleman.cpp) and the Python restatement oracle/pyref_needleman.py against alignments the
reference's own class files produced (oracle/make_ref_needleman.py -> tests/golden/ref_needleman.npz) and against the
nMismatchDiffBestvsSecondBest / MORE_THAN_ONE_MATCH outcomes of the findUMI vectors.  Host arithmetic: no GPU needed (the records of the
findUMI test come from the oracle here and from the GPU kernel in the -m gpu twin)."""
import os

import numpy as np
import pytest

import __graft_entry__ as g
import workloads
from oracle import pyref_needleman as P
from test_ref_vectors import FIND_UMI_FILES, GOLDEN, TEST_BARCODES_FILES, _find_umi_inputs, _test_barcodes_inputs

KEYS = ("leading_gap_1", "leading_gap_2", "trailing_gap_1", "trailing_gap_2", "indel", "mismatch", "match")


@pytest.fixture(scope="module")
def pk():
    p = g.load_package()
    p.build()
    return p


def test_needleman_matches_reference_bytecode(pk):
    z = np.load(os.path.join(GOLDEN, "ref_needleman.npz"))
    assert len(z["L"]) >= 600 and z["custom"].sum() > 100
    hist = np.zeros(16, dtype=int)
    for i in range(len(z["L"])):
        a, b, L = str(z["template"][i]), str(z["read"][i]), int(z["L"][i])
        sc = dict(zip(KEYS, (int(x) for x in z["scores"][i]))) if z["custom"][i] else None
        want = tuple(int(x) for x in z["counts"][i])
        al = P.align(a, b, sc)
        assert al == (str(z["match"][i]), str(z["pattern"][i]), str(z["read_row"][i])), i          # the alignment strings themselves
        assert P.count_errors(*al) == want, i
        got = pk.needleman_errors(workloads.g_pack(a), workloads.g_pack(b), L, pk.NeedlemanScores(**sc) if sc else None)      # slr_needleman_errors
        assert got == want, (i, a, b, got, want)
        hist[min(want[3], 15)] += 1
    assert (hist[:6] > 10).all()
    assert (z["counts"][:, 0] > 0).sum() > 50 and (z["counts"][:, 1] > 0).sum() > 50             # insertions and deletions occur


def _diff_check(pk, res_of):
    n_second = n_tie = 0
    for fname in FIND_UMI_FILES:
        z = np.load(os.path.join(GOLDEN, fname))
        for i in range(len(z["ed"])):
            if str(z["exc"][i]):
                continue
            sl, anchor, ed, pm, post_len, slen = _find_umi_inputs(z, i)
            umis = z["umis"][z["umi_offsets"][i]:z["umi_offsets"][i + 1]]
            res = res_of(umis, sl, anchor, ed, pm, post_len, int(z["bail"][i]), slen)
            d = pk.guided_mismatch_diff(res, sl, anchor, 12, slice_len=slen)
            row = z["row"][i]
            if int(row[7]):                                                   # the reference built a second-best Match
                assert int(d[0]) == int(row[10]), (fname, i, d, row)          # nMismatchDiffBestvsSecondBest
                assert (int(d[0]) == 0) == bool(int(row[1]) & 4), (fname, i)  # MORE_THAN_ONE_MATCH <=> diff == 0
                assert bool(row[0]) == (int(d[0]) != 0), (fname, i)           # found <=> not tied (list not empty here)
                n_second += 1
                n_tie += int(d[0]) == 0
            else:
                assert int(d[0]) == pk.G_NO_SECOND
    assert n_second >= 80 and n_tie >= 9


def test_mismatch_diff_matches_find_umi_bytecode(pk, orc):
    """records from the CPU oracle, the Needleman step from the library: diff, flag and `found` as IlluminaUMIanalyzer.findUMI computed them"""
    def res_of(umis, sl, anchor, ed, pm, post_len, bail, slen):
        return orc.guided_batch(umis, np.array([0, len(umis)], dtype=np.int64), sl, anchor, np.array([0], dtype=np.int32), ed, 12, pm, post_len,
                                bailout=bail, slice_len=slen)[0]
    _diff_check(pk, res_of)


@pytest.mark.gpu
def test_gpu_mismatch_diff_matches_find_umi_bytecode(pk, ctx):
    """the same with the records of the guided kernel: the S4 seam now answers found / not found like findUMI"""
    def res_of(umis, sl, anchor, ed, pm, post_len, bail, slen):
        return pk.GuidedSets(ctx, umis, np.array([0, len(umis)], dtype=np.int64), 12).match(sl, anchor, np.array([0], dtype=np.int32), ed, pm, post_len,
                                                                                            bailout=None if bail < 0 else bail, slice_len=slen)[0]
    _diff_check(pk, res_of)


def _bc_diff_check(pk, res_of):
    n_second = n_tie = 0
    for fname in TEST_BARCODES_FILES:
        z = np.load(os.path.join(GOLDEN, fname))
        for i in range(len(z["ed"])):
            if str(z["exc"][i]):
                continue
            sl, anchor, slen, gene, allk, empk = _test_barcodes_inputs(z, i)
            res = res_of(gene, allk, empk, sl, anchor, int(z["ed"][i]), int(z["pm"][i]), int(z["bail"][i]), slen)
            d = int(pk.guided_mismatch_diff(res, sl, anchor, 16, slice_len=slen)[0])
            if z["n_distinct"][i] == 2:
                assert d == int(z["mismatch_diff"][i]), (fname, i, d, z["mismatch_diff"][i])
                assert (d == 0) == bool(int(z["bc_flag"][i]) & 64), (fname, i)                     # BarcodeFindingFlag.MORE_THAN_ONE_MATCH
                n_second += 1
                n_tie += d == 0
            else:
                assert d == pk.G_NO_SECOND and not int(z["bc_flag"][i]) & 64
    assert n_second >= 60 and n_tie >= 10


def test_mismatch_diff_matches_test_barcodes_bytecode(pk, orc):
    """BC flavour (getBestAndSecondBCorUMI(CELLBC) after testBarcodes): records from the CPU oracle, the Needleman step from the library"""
    def res_of(gene, allk, empk, sl, anchor, ed, pm, bail, slen):
        return orc.guided_batch(gene, np.array([0, len(gene)], dtype=np.int64), sl, anchor, np.array([0], dtype=np.int32), ed, 16, pm, 10, bailout=bail,
                                bc_flavour=True, all_keys=allk, all_ed=3, empty_keys=empk, empty_ed=2, slice_len=slen)[0]
    _bc_diff_check(pk, res_of)



def test_mismatch_diff_refuses_foreign_records(pk):
    res = np.zeros(1, dtype=pk.GUIDED_RESULT)
    res["n_distinct"] = 2
    res["offset"][0] = [30, 0]
    sl = np.frombuffer(b"A" * 32, dtype=np.uint8).reshape(1, 32)
    with pytest.raises(pk.SiceloreGpuError) as e:
        pk.guided_mismatch_diff(res, sl, np.array([5], dtype=np.int32), 12)
    assert e.value.code == pk.SLR_E_INVALID
    with pytest.raises(pk.SiceloreGpuError):
        pk.needleman_errors(0, 0, 33)
