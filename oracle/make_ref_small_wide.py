"""Wider bytecode pins of three small pieces (generators and layouts of oracle/make_ref_vectors.py, other seeds): LevenshteinDistance.limitedCompare
(4 000 pairs), the best-of-9 packing of ClusteringEditDistanceBase (1 500 pairs) and pass 1's per-read lambda (2 x 600 reads).

    python oracle/make_ref_small_wide.py      -> tests/golden/ref_levenshtein_wide.npz, ref_best9_wide.npz, ref_exact_lookup_wide.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_ref_vectors as M  # noqa: E402
from oracle import minijvm as J  # noqa: E402


def main():
    vm = J.VM(M.JARS + [M.REF + "/lib/commons-lang3-3.17.0.jar"])
    lv = M.lev_cases(vm, np.random.default_rng(9400), 4000)
    np.savez_compressed(os.path.join(M.OUT, "ref_levenshtein_wide.npz"), rows=lv)
    print("levenshtein", lv.shape, "d histogram", np.bincount(lv[:, 24].astype(np.int8).astype(int) + 1))
    b9, eq = M.best9_cases(vm, np.random.default_rng(9499), 1500)
    np.savez_compressed(os.path.join(M.OUT, "ref_best9_wide.npz"), rows=b9, equality=np.int64(eq))
    print("best-of-9", b9.shape, "ED histogram", np.bincount(b9[:, 28] & 0xFFFFFF))
    n = 600
    ex = M.exact_lookup_cases(vm, np.random.default_rng(9408), n)
    np.savez_compressed(os.path.join(M.OUT, "ref_exact_lookup_wide.npz"), three_prime=np.array([c["tp"] for c in ex], dtype=np.int32),
                        whitelist=np.array([c["whitelist"] for c in ex], dtype=np.uint64), read=np.array([[r[0] for r in c["reads"]] for c in ex]),
                        adapterpos=np.array([[r[1] for r in c["reads"]] for c in ex], dtype=np.int32),
                        found=np.array([[r[2] for r in c["reads"]] for c in ex], dtype=np.int32),
                        count_keys=np.array([sorted(c["counts"]) + [0] * (n - len(c["counts"])) for c in ex], dtype=np.uint64),
                        count_vals=np.array([[c["counts"][k] for k in sorted(c["counts"])] + [0] * (n - len(c["counts"])) for c in ex], dtype=np.int64))
    print("pass-1 exact lookup", [(sum(r[2] == 1 for r in c["reads"]), sum(r[2] == -1 for r in c["reads"])) for c in ex], "(found, throwing) per geometry")


if __name__ == "__main__":
    main()
