"""Wider bytecode pin of the Illumina-guided search: the generators of oracle/make_ref_vectors.py (findUMI as a whole,
testBarcodes + getBestAndSecondBCorUMI, matchesSeqEditDistance windows) run on more inputs with other seeds, fanned out over the cores (one
interpreter per worker).  Same file layout as the narrow sets, so the same tests read both:

    python oracle/make_ref_guided_wide.py [n_find_umi n_test_barcodes n_windows]      # default 192 96 160
    -> tests/golden/ref_find_umi_wide.npz, ref_test_barcodes_wide.npz, ref_guided_wide.npz
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_ref_vectors as M  # noqa: E402
from oracle import minijvm as J  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def work(job):
    kind, seed, n = job
    vm = J.VM(M.JARS + [M.REF + "/lib/commons-lang3-3.17.0.jar"])
    rng = np.random.default_rng(seed)
    t0 = time.time()
    if kind == "find_umi":
        out = M.find_umi_cases(vm, rng, n)
    elif kind == "test_barcodes":
        out = M.test_barcodes_cases(vm, rng, n)
    else:
        out = M.guided_cases(vm, rng, n)
    print("  %s seed %d: %d cases, %.0f s, %d bytecodes" % (kind, seed, len(out), time.time() - t0, vm.n_insn), flush=True)
    return kind, seed, out


def main():
    n_fu, n_tb, n_gc = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (192, 96, 160)
    per = 8
    jobs = [("test_barcodes", 9100 + k, per) for k in range(n_tb // per)] + [("find_umi", 6100 + k, per) for k in range(n_fu // per)] + \
           [("guided", 3100 + k, 2 * per) for k in range(n_gc // (2 * per))]
    t0 = time.time()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(work, jobs, chunksize=1)
    by = {"find_umi": [], "test_barcodes": [], "guided": []}
    for kind, seed, out in sorted(res, key=lambda r: (r[0], r[1])):
        by[kind] += out
    fu, tb, gc = by["find_umi"], by["test_barcodes"], by["guided"]
    flat = M.flat
    keys, koff = flat(fu, "umis")
    np.savez_compressed(os.path.join(OUT, "ref_find_umi_wide.npz"), stranded=np.array([c["stranded"] for c in fu]), bc_end=np.array([c["bc_end"] for c in fu], dtype=np.int32),
                        umis=keys, umi_offsets=koff, ed=np.array([c["ed"] for c in fu], dtype=np.int32), pm=np.array([c["pm"] for c in fu], dtype=np.int32),
                        bail=np.array([c["bail"] for c in fu], dtype=np.int32), exc=np.array([c["exc"] for c in fu]),
                        row=np.array([c["row"][:12] for c in fu], dtype=np.int64), second=np.array([c.get("second", "") for c in fu]))
    print("findUMI", len(fu), "reads, found", sum(c["row"][0] for c in fu), "with second", sum(c["row"][7] for c in fu if not c["exc"]), "throwing", sum(bool(c["exc"]) for c in fu))
    old = M.OUT
    try:                                                                                      # save_test_barcodes writes OUT/ref_test_barcodes.npz
        tmp = os.path.join(OUT, "_wide_tmp")
        os.makedirs(tmp, exist_ok=True)
        M.OUT = tmp
        M.save_test_barcodes(tb)
        os.replace(os.path.join(tmp, "ref_test_barcodes.npz"), os.path.join(OUT, "ref_test_barcodes_wide.npz"))
        os.rmdir(tmp)
    finally:
        M.OUT = old
    print("testBarcodes", len(tb), "reads, with hits", sum(c["n_raw"] > 0 for c in tb), "with second", sum(c["n_distinct"] == 2 for c in tb))
    keys, koff = flat(gc, "keys")
    ak, aoff = flat(gc, "allk")
    ek, eoff = flat(gc, "empk")
    res = np.array([(i,) + r for i, c in enumerate(gc) for r in c["res"]], dtype=np.int64).reshape(-1, 7)
    np.savez_compressed(os.path.join(OUT, "ref_guided_wide.npz"), keys=keys, key_offsets=koff, all_keys=ak, all_offsets=aoff, empty_keys=ek, empty_offsets=eoff,
                        w=np.array([c["w"] for c in gc], dtype=np.uint64), L=np.array([c["L"] for c in gc], dtype=np.int32),
                        ed=np.array([c["ed"] for c in gc], dtype=np.int32), bc=np.array([c["bc"] for c in gc], dtype=np.int32),
                        bail=np.array([c["bail"] for c in gc], dtype=np.int32), off=np.array([c["off"] for c in gc], dtype=np.int32),
                        post=np.array([c["post"].ljust(10, "-") for c in gc]), res=res,
                        flag_gene=np.int64(512), flag_all=np.int64(4), flag_empty=np.int64(8))
    print("guided", len(gc), "windows,", len(res), "list entries; total %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
