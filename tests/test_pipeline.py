"""scanfastq from pass 1 to pass 2 without a JVM (README.md:214-226 without `-g`: the used-barcode list is discovered, not given): exact lookup of
every read against the 10x whitelist -> unfilteredUsedBarcodeMap -> count filter -> collision test -> collision merge + ranks -> pass-2
assignment against the used list.  The chain is written once and run with two back ends: the CPU oracle (here) and the library on the GPU
(-m gpu), whose records must be identical; the host steps between the passes are the library's in both (they need no device)."""
import numpy as np
import pytest

import __graft_entry__ as g


class OracleBackend:
    def __init__(self, orc):
        self.orc = orc

    def exact(self, whitelist, slices, anchor):
        rec = self.orc.exact_lookup_batch(self.orc.BarcodeSet(whitelist, np.arange(1, len(whitelist) + 1, dtype=np.int32)), slices, anchor)
        hit = rec["flags"] & 1 != 0
        cnt = np.zeros(len(whitelist), dtype=np.int64)
        np.add.at(cnt, np.searchsorted(whitelist, rec["bc"][hit]), 1)                       # whitelist is sorted
        return rec, cnt

    def collide(self, barcodes, ed):
        return self.orc.collide_batch(self.orc.BarcodeSet(barcodes), barcodes, ed)[0]

    def assign(self, barcodes, rank, slices, anchor, ed):
        return self.orc.assign_barcode_batch(self.orc.BarcodeSet(barcodes, rank), slices, anchor, ed)[0]


class GpuBackend:
    def __init__(self, pkg, ctx):
        self.pkg, self.ctx = pkg, ctx

    def exact(self, whitelist, slices, anchor):
        table = self.pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(self.ctx, whitelist)
        rec = self.pkg.UsedCellBCListGenerator(self.ctx, table).addFastqs(slices, anchor)
        return rec, table.counts()[:, 0]

    def collide(self, barcodes, ed):
        table = self.pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(self.ctx, barcodes)
        return self.pkg.BarcodeDatasetColissionTester(self.ctx, table, ed).colissionsFromScan(barcodes)

    def assign(self, barcodes, rank, slices, anchor, ed):
        table = self.pkg.BarcodesMapForBCfinding(self.ctx, barcodes, rank)
        return self.pkg.Parser(self.ctx, table, bcEditDistance=ed).assign_barcodes(slices, anchor)


def run_pipeline(pkg, be, ed=1, n_reads=60000, n_cells=300):
    wl = np.sort(pkg.synth_whitelist(20000, 77))
    slices, anchor, truth = pkg.synth_reads(wl, n_reads, seed=21, n_cells=n_cells)          # reads of n_cells cells (skewed depths) + noise
    # ---- pass 1 (UsedCellBCListGenerator$Worker): exact lookup, unfilteredUsedBarcodeMap
    rec1, counts = be.exact(wl, slices, anchor)
    used = np.nonzero(counts)[0]
    # ---- finalizeData: count filter, collision tester, merge, ranks (host steps of the library)
    f = pkg.used_filter_low_counts(counts[used], n_reads)
    lst, lst_counts = wl[used][f], counts[used][f].astype(np.int32)
    col = be.collide(lst, ed)
    keep, rank, flags = pkg.used_merge_collisions(lst, lst_counts, col, 10, ed, 500)
    final, final_rank = lst[keep], rank[keep]
    # ---- pass 2 (Parser.assignBarcode against the used list)
    rec2 = be.assign(final, final_rank, slices, anchor, ed)
    return dict(wl=wl, truth=truth, rec1=rec1, counts=counts, lst=lst, col=col, keep=keep, rank=rank, flags=flags, final=final, rec2=rec2)


def check_sanity(r, n_cells):
    assert 0.5 * n_cells < len(r["final"]) <= 1.2 * n_cells                                  # the list is about the cells that were sequenced
    assert len(r["lst"]) >= len(r["final"]) and not r["flags"] & 1
    ok = r["rec2"]["flags"] & 1 != 0
    assert ok.mean() > 0.5
    assert set(r["rec2"]["bc"][ok].tolist()) <= set(r["final"].tolist())
    assert (r["rec2"]["rank"][ok] >= 1).all() and r["rec2"]["rank"][ok].max() <= len(r["final"])
    assert ok.sum() > (r["rec1"]["flags"] & 1 != 0).sum() * 0.9                              # ED 1 against the short list rescues at least what pass 1 saw exactly


def test_pipeline_on_the_oracle(orc):
    pkg = g.load_package()
    pkg.build()
    r = run_pipeline(pkg, OracleBackend(orc))
    check_sanity(r, 300)
