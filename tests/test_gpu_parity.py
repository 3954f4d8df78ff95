"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs;
golden fixtures; size-independent properties at BASELINE.json's full sizes."""
import glob
import os

import numpy as np
import pytest

import workloads

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_assign(pkg, ctx, wl, rank, slices, anchors, ed, pm, three_prime, lens=None):
    table = pkg.BarcodesMapForBCfinding(ctx, wl, rank)
    res = pkg.Parser(ctx, table, ed, pm, three_prime).assign_barcodes(slices, anchors, lens=lens)
    return res, table


@pytest.mark.parametrize("three_prime", [True, False])
@pytest.mark.parametrize("ed", [0, 1, 2])
@pytest.mark.parametrize("skew", [False, True])
def test_bc_adversarial(pkg, orc, ctx, three_prime, ed, skew):
    _, slices, anchors, wl = workloads.adversarial(5000 + ed + 10 * skew, three_prime, 600, skew=skew)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, three_prime)
    got, table = gpu_assign(pkg, ctx, wl, rank, slices, anchors, ed, 2, three_prime)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    # BarcodesAssigned.tsv counters (Parser.java:L305-L311)
    counts = table.counts()
    ok = (exp["flags"] & 1) == 1
    expc = np.zeros_like(counts)
    idx = {int(k): i for i, k in enumerate(wl)}
    for r in exp[ok]:
        expc[idx[int(r["bc"])], r["ed"]] += 1
    assert (counts == expc).all()


@pytest.mark.parametrize("three_prime", [True, False])
@pytest.mark.parametrize("ed", [1, 2])
def test_bc_dense_overflow(pkg, orc, ctx, three_prime, ed):
    _, slices, anchors, wl = workloads.adversarial(6000 + ed, three_prime, 80, skew=True, dense=True, nrand=50)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), slices, anchors, ed, 2, three_prime)
    got, _ = gpu_assign(pkg, ctx, wl, None, slices, anchors, ed, 2, three_prime)
    assert (got == exp).all()


@pytest.mark.parametrize("pm", [0, 1, 3, 4])
def test_bc_plusminus(pkg, orc, ctx, pm):
    _, slices, anchors, wl = workloads.adversarial(7000 + pm, True, 200, anchor=10)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), slices, anchors, 2, pm, True)
    got, _ = gpu_assign(pkg, ctx, wl, None, slices, anchors, 2, pm, True)
    assert (got == exp).all()


def test_bc_ragged_empty_and_errors(pkg, orc, ctx):
    _, slices, anchors, wl = workloads.adversarial(78, True, 64)
    lens = np.random.default_rng(1).integers(0, 33, size=64).astype(np.int32)
    bs = orc.BarcodeSet(wl)
    exp = np.zeros(64, dtype=orc.BC_RESULT)
    for i in range(64):
        exp[i] = orc.assign_barcode_batch(bs, slices[i:i + 1], anchors[i:i + 1], 2, 2, True, slice_len=int(lens[i]))[0][0]
    got, table = gpu_assign(pkg, ctx, wl, None, slices, anchors, 2, 2, True, lens=lens)
    assert (got == exp).all()
    p = pkg.Parser(ctx, table, 2)
    assert len(p.assign_barcodes(slices[:0], anchors[:0])) == 0
    with pytest.raises(pkg.SiceloreGpuError) as e:
        pkg.Parser(ctx, table, 3).assign_barcodes(slices, anchors)        # --bcEditDistance 3: refused, not approximated
    assert e.value.code == pkg.SLR_E_UNSUPPORTED
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.BarcodesMapForBCfinding(ctx, wl, None, bc_len=14)


def test_bc_golden(pkg, ctx):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "bc_*.npz"))):
        g = np.load(f)
        got, _ = gpu_assign(pkg, ctx, g["whitelist"], g["rank"], g["slices"], g["anchor"], int(g["ed"]), 2, bool(g["three_prime"]))
        assert (got == g["result"]).all(), f


@pytest.mark.parametrize("cfg", [("737k", 737280, 737, 1, 1, 60000), ("737k", 737280, 737, 1, 2, 6000),
                                 ("3m", 3000000, 3000000, 2, 2, 6000), ("3m", 3000000, 3000000, 2, 1, 60000)])
def test_bc_synthetic_configs(pkg, orc, ctx, cfg):
    """configs[1] / configs[2] of BASELINE.json at a size the oracle finishes in seconds"""
    _, n_wl, wl_seed, read_seed, ed, n = cfg
    wl = pkg.synth_whitelist(n_wl, wl_seed)
    slices, anchors, _ = pkg.synth_reads(wl, n, seed=read_seed)
    rank = np.arange(1, n_wl + 1, dtype=np.int32)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, True)
    got, _ = gpu_assign(pkg, ctx, wl, rank, slices, anchors, ed, 2, True)
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    assert ((got["flags"] & 1) == 1).mean() > 0.5


def test_bc_five_prime_synthetic(pkg, orc, ctx):
    wl = pkg.synth_whitelist(100000, 5)
    slices, anchors, _ = pkg.synth_reads(wl, 8000, seed=6, three_prime=False)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl), slices, anchors, 2, 2, False)
    got, _ = gpu_assign(pkg, ctx, wl, None, slices, anchors, 2, 2, False)
    assert (got == exp).all()
    assert ((got["flags"] & 1) == 1).mean() > 0.5


def test_bc_full_size_properties(pkg, orc, ctx):
    """configs[2] at full size (10 M reads, 3 M list, ED 2): determinism, chunk invariance, membership, counters, and a
    bit-exact spot check of a strided sample against the oracle."""
    import torch
    n = 10_000_000
    wl = pkg.synth_whitelist(3_000_000, 3_000_000)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    slices, anchors, truth = pkg.synth_reads(wl, n, seed=2)
    table = pkg.BarcodesMapForBCfinding(ctx, wl, rank)
    parser = pkg.Parser(ctx, table, 2)
    a = parser.assign_barcodes(slices, anchors)
    # device-pointer entry point on torch-owned memory gives the same bytes
    d_sl = torch.from_numpy(slices).cuda()
    d_an = torch.from_numpy(anchors).cuda()
    d_out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    table.reset_counts()
    parser.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    b = d_out.cpu().numpy().view(pkg.BC_RESULT).reshape(-1)
    assert (a == b).all()
    ok = (a["flags"] & 1) == 1
    assert 0.6 < ok.mean() < 0.8
    assert (a["flags"] & ~np.uint32(1)).max() == 0
    # every assigned barcode is a list member with the right rank; ED bounds; second-best rule
    srt = np.sort(wl)
    pos = np.searchsorted(srt, a["bc"][ok])
    assert (srt[np.minimum(pos, len(srt) - 1)] == a["bc"][ok]).all()
    order = np.argsort(wl)
    assert (rank[order][pos] == a["rank"][ok]).all()
    assert ((a["ed"][ok] >= 0) & (a["ed"][ok] <= 2) & (a["ed"][ok] < a["ed_second"][ok])).all()
    assert ((a["ed"][~ok] == -1) | (a["ed_second"][~ok] <= a["ed"][~ok]) | (a["ed"][~ok] > 2)).all()
    # counters = histogram of the assigned reads (run once since the reset)
    counts = table.counts()
    assert counts.sum() == ok.sum()
    assert (counts.sum(axis=0) == np.bincount(a["ed"][ok], minlength=3)).all()
    # reads generated from a list barcode and assigned: overwhelmingly the true one
    t = ok & (truth >= 0)
    assert (a["bc"][t] == wl[truth[t]]).mean() > 0.85       # ED 2 on a dense list trades accuracy for yield (README.md:100,180)
    # bit-exact vs oracle on a strided sample
    sel = np.arange(0, n, n // 4000)
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices[sel], anchors[sel], 2)
    assert (a[sel] == exp).all()


# ---- UMI ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("umi_len", [12, 10, 14, 8])
def test_umi_parity(pkg, orc, ctx, umi_len):
    umis, offs = workloads.umi_jobs(140 + umi_len, umi_len, n_jobs=80, max_n=70)
    exp, oo = orc.umi_matrix_batch(umis, offs, umi_len)
    got, oo2 = pkg.generate_distance_matrices(ctx, umis, offs, umi_len)
    assert (oo == oo2).all() and (got == exp).all()


def test_umi_golden_and_edge_cases(pkg, orc, ctx):
    for f in sorted(glob.glob(os.path.join(GOLDEN, "umi_*.npz"))):
        g = np.load(f)
        got, _ = pkg.generate_distance_matrices(ctx, g["umis"], g["job_offsets"], int(g["umi_len"]))
        assert (got == g["matrix"]).all()
    # empty job list, single-read jobs, non-contiguous out_offsets
    got, oo = pkg.generate_distance_matrices(ctx, np.zeros((0, 16), np.uint8), np.zeros(1, np.int64))
    assert len(got) == 0
    umis, offs = workloads.umi_jobs(9, 12, n_jobs=5, max_n=6)
    sizes = np.diff(offs)
    oo = np.concatenate([[0], np.cumsum(sizes * sizes + 3)]).astype(np.int64)      # 3 cells of padding after each job
    out = np.full(int(oo[-1]), -7, dtype=np.int32)
    pkg.generate_distance_matrices(ctx, umis, offs, 12, out=out, out_offsets=oo)
    exp, eo = orc.umi_matrix_batch(umis, offs, 12)
    for j in range(5):
        n2 = int(sizes[j]) ** 2
        assert (out[oo[j]:oo[j] + n2] == exp[eo[j]:eo[j] + n2]).all()
        assert (out[oo[j] + n2:oo[j + 1]] == -7).all()
    with pytest.raises(pkg.SiceloreGpuError):
        pkg.generate_distance_matrices(ctx, umis, offs, 15)


def test_umi_config4_sample_and_large_job(pkg, orc, ctx):
    """configs[3]-shaped input (many small (cell, gene) jobs) vs the oracle, plus one deep job checked by properties"""
    umis, offs = pkg.synth_umi_jobs(40000, mean=4.0, cap=2000, seed=4)
    got, oo = pkg.generate_distance_matrices(ctx, umis, offs)
    exp, _ = orc.umi_matrix_batch(umis, offs)
    assert (got == exp).all()
    umis, offs = pkg.synth_umi_jobs(1, mean=1e9, cap=6000, seed=8)
    n = int(offs[1])
    assert n == 6000
    got, _ = pkg.generate_distance_matrices(ctx, umis, offs)
    m = got.reshape(n, n)
    B = pkg.BestEditDistance
    assert (np.diag(m) == (0 | (0x08000000 << 1) | (0x01000000 << 1))).all()
    assert (B.getED(m) == B.getED(m.T)).all() and B.getED(m).max() <= 5
    assert (B.getPos1(m) == B.getPos2(m.T)).all()
    rows = np.arange(0, n, 97)
    exp_rows, _ = orc.umi_matrix_batch(umis, offs)                                   # 36 M cells on the CPU: a few seconds
    assert (m[rows] == exp_rows.reshape(n, n)[rows]).all()


def test_concurrent_callers_share_one_context(pkg, orc):
    """The reference's nCPU Parser workers call into the seam concurrently (WorkerReadscanner.java:L186-L204): many host threads,
    one context with fewer stream slots than threads, all four entry points interleaved — every call returns what a serial
    call returns, and the table's counters add up."""
    from concurrent.futures import ThreadPoolExecutor
    ctx = pkg.Context(0, n_streams=2)
    wl = pkg.synth_whitelist(200000, 21)
    rank = np.arange(1, len(wl) + 1, dtype=np.int32)
    table = pkg.BarcodesMapForBCfinding(ctx, wl, rank)
    used = np.unique(np.concatenate([wl[:20000], wl[:10000] ^ np.uint64(1 << 9)]))
    utable = pkg.BarcodesMapForBCfinding(ctx, used)
    parser = pkg.Parser(ctx, table, 2)
    batches = [pkg.synth_reads(wl, 30000 + 1000 * i, seed=40 + i)[:2] for i in range(6)]
    umi_jobs = [pkg.synth_umi_jobs(3000 + 100 * i, mean=4.0, cap=300, seed=60 + i) for i in range(6)]
    serial_bc = [parser.assign_barcodes(s, a) for s, a in batches]
    serial_umi = [pkg.generate_distance_matrices(ctx, u, o)[0] for u, o in umi_jobs]
    serial_col = pkg.BarcodeDatasetColissionTester(ctx, utable, 2).colissionsFromScan()
    table.reset_counts()

    def work(i):
        kind = i % 3
        if kind == 0:
            return ("bc", i // 3 % 6, parser.assign_barcodes(*batches[i // 3 % 6]))
        if kind == 1:
            return ("umi", i // 3 % 6, pkg.generate_distance_matrices(ctx, *umi_jobs[i // 3 % 6])[0])
        return ("col", 0, pkg.BarcodeDatasetColissionTester(ctx, utable, 2).colissionsFromScan())

    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(work, range(36)))
    n_bc = 0
    for kind, j, r in results:
        if kind == "bc":
            assert (r == serial_bc[j]).all()
            n_bc += int((r["flags"] & 1).sum())
        elif kind == "umi":
            assert (r == serial_umi[j]).all()
        else:
            assert (r == serial_col).all()
    assert table.counts().sum() == n_bc                      # device-side atomics: nothing lost under concurrency
    # and the oracle agrees with one of the batches
    exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), batches[0][0][:3000], batches[0][1][:3000], 2)
    assert (serial_bc[0][:3000] == exp).all()
