"""Wider bytecode pins of the two inner loops (same generators and file layouts as oracle/make_ref_vectors.py, other seeds, fanned out over the
cores): calcEditDistances (ClusteringEditDistanceBase.lambda$static$7) on 2 000 read pairs and BarcodeMatchTester.doJob on 240 windows.

    python oracle/make_ref_wide2.py [n_pairs n_windows]      -> tests/golden/ref_umi_pairs_wide.npz, ref_dojob_wide.npz
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
    out = M.umi_pair_cases(vm, rng, n) if kind == "pairs" else M.dojob_cases(vm, rng, n)
    print("  %s seed %d: %d cases, %.0f s" % (kind, seed, len(out), time.time() - t0), flush=True)
    return kind, seed, out


def main():
    n_pairs, n_win = [int(x) for x in sys.argv[1:3]] if len(sys.argv) > 2 else (2000, 240)
    jobs = [("dojob", 5200 + k, 10) for k in range(n_win // 10)] + [("pairs", 4300 + k, 250) for k in range(n_pairs // 250)]
    t0 = time.time()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(work, jobs, chunksize=1)
    up = [c for k, s, out in sorted(res, key=lambda r: (r[0], r[1])) if k == "pairs" for c in out]
    dj = [c for k, s, out in sorted(res, key=lambda r: (r[0], r[1])) if k == "dojob" for c in out]
    np.savez_compressed(os.path.join(OUT, "ref_umi_pairs_wide.npz"), s1=np.array([r[0] for r in up]), s2=np.array([r[1] for r in up]),
                        end1=np.array([r[2] for r in up], dtype=np.int32), end2=np.array([r[3] for r in up], dtype=np.int32),
                        five_prime=np.array([r[4] for r in up], dtype=np.int32), packed=np.array([r[5] for r in up], dtype=np.int64))
    print("calcEditDistances", len(up), "pairs, ED histogram", np.bincount(np.array([r[5] for r in up]) & 0xFFFFFF))
    keys, koff = M.flat(dj, "keys")
    resrows = np.array([(i,) + r for i, c in enumerate(dj) for r in c["res"]], dtype=np.int64).reshape(-1, 8)
    np.savez_compressed(os.path.join(OUT, "ref_dojob_wide.npz"), keys=keys, key_offsets=koff, w=np.array([c["w"] for c in dj], dtype=np.uint64),
                        ed=np.array([c["ed"] for c in dj], dtype=np.int32), mode=np.array([c["mode"] for c in dj], dtype=np.int32),
                        post=np.array([c["post"].ljust(5, "-") for c in dj]), off=np.array([c["off"] for c in dj], dtype=np.int32), res=resrows)
    print("doJob", len(dj), "windows,", len(resrows), "matches; total %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
