"""Wider bytecode pin of ClusterOne_MyClustering.call (the clusterer of the jobs above 100 reads): the generator of oracle/make_ref_myclust.py
(same injections and shims, same file layout) on more jobs with other seeds, one interpreter per core.

    python oracle/make_ref_myclust_wide.py [n_jobs]      -> tests/golden/ref_myclust_wide.npz
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_ref_hier as H  # noqa: E402
from oracle import make_ref_myclust as MC  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_myclust_wide.npz")


def work(job):
    seed, n_jobs = job
    H.install_set_extras()
    vm = MC.MVM(H.JARS)
    H.install_overrides(vm)
    rng = np.random.default_rng(seed)
    t0 = time.time()
    jobs = []
    for t in range(n_jobs):
        n = int(rng.integers(101, 150)) if t % 3 else int(rng.integers(150, 280))
        if t == 0:
            n = int(rng.integers(20, 60))                       # the class itself has no size limit
        packed = H.random_packed(rng, n, t % 3) if t % 5 else MC.umi_like(rng, n)
        qv = int(rng.integers(0, 2))
        prm = (2, 1, 3000, int(rng.choice([50, 50, 3])))
        res, n_found = H.run_job(vm, packed.tolist(), prm, qv, cls="ClusterOne_MyClustering")
        jobs.append(dict(n=n, packed=packed, qv=qv, prm=prm, res=res, n_found=n_found))
    print("  seed %d: %d jobs, %.0f s, %d bytecodes" % (seed, n_jobs, time.time() - t0, vm.n_insn), flush=True)
    return seed, jobs


def main():
    n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    per = 6
    t0 = time.time()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(work, [(88000 + k, per) for k in range(n_jobs // per)], chunksize=1)
    jobs = [j for _, js in sorted(res, key=lambda r: r[0]) for j in js]
    off = np.cumsum([0] + [j["n"] for j in jobs]).astype(np.int64)
    moff = np.cumsum([0] + [j["n"] ** 2 for j in jobs]).astype(np.int64)
    flat = lambda k, dt: np.array([r[k] for j in jobs for r in j["res"]], dtype=dt)
    np.savez_compressed(OUT, job_offsets=off, out_offsets=moff, packed=np.concatenate([j["packed"].ravel() for j in jobs]),
                        qv01=np.array([j["qv"] for j in jobs], dtype=np.uint8), params=np.array([j["prm"] for j in jobs], dtype=np.int32),
                        assigned=flat("assigned", np.int8), u8=np.array([r["u8"] for j in jobs for r in j["res"]]), u1=flat("u1", np.int8),
                        u2=flat("u2", np.int8), pos2=flat("pos2", np.int8), flagval=flat("flagval", np.int64),
                        n_found=np.array([j["n_found"] for j in jobs], dtype=np.int32))
    print("ClusterOne_MyClustering.call (wide): %d jobs, %d reads, %d assigned, %.0f s" % (len(jobs), int(off[-1]), int(flat("assigned", np.int8).sum()), time.time() - t0))


if __name__ == "__main__":
    main()
