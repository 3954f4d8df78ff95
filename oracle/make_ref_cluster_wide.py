"""Wider bytecode pin of ClusterOne_MyClustering.clusterLocal (seam S5): the generator of oracle/make_ref_vectors.py (cluster_cases, same layout)
on 480 more jobs with another seed.

    python oracle/make_ref_cluster_wide.py [n_jobs]      -> tests/golden/ref_cluster_local_wide.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_ref_vectors as M  # noqa: E402
from oracle import minijvm as J  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 480
    vm = J.VM(M.JARS + [M.REF + "/lib/commons-lang3-3.17.0.jar"])
    t0 = time.time()
    cc = M.cluster_cases(vm, np.random.default_rng(272727), n)
    M.save_cluster_cases(cc)                                                   # writes OUT/ref_cluster_local.npz: redirected below
    print("clusterLocal (wide):", len(cc), "jobs, with clusters", sum(c["present"] for c in cc), "reads in clusters",
          sum(int((c["label"] >= 0).sum()) for c in cc), "%.0f s" % (time.time() - t0), vm.n_insn, "bytecodes")


if __name__ == "__main__":
    tmp = os.path.join(M.OUT, "_cluster_wide_tmp")
    os.makedirs(tmp, exist_ok=True)
    old = M.OUT
    M.OUT = tmp
    try:
        main()
        os.replace(os.path.join(tmp, "ref_cluster_local.npz"), os.path.join(old, "ref_cluster_local_wide.npz"))
    finally:
        M.OUT = old
        os.rmdir(tmp)
