"""Golden vectors of the large-job clusterer FROM THE REFERENCE'S OWN CLASS FILES: ClusterOne_MyClustering.call
(F!com/rw/umifinder/analyzers/clustering/ClusterOne_MyClustering.class, ClusterOne_MyClustering.java:L59-L219) run by oracle/minijvm.py together
with OneUmiCluster (setClusterCenter, removeEntries), DistanceMatrix, ClusterOneBase.setSamflagsAndStatsForClustered / flagDontUMIassignRecords,
BestEditDistance, PlusMinusOneEnum and commons-lang3's ImmutablePair as bytecode.  Frozen in tests/golden/ref_myclust.npz.

    python oracle/make_ref_myclust.py [n_jobs]

Same injections as make_ref_hier.py (the packed matrix, the SAM side of OneNanoporeResult).  The class's streams are parallel above 30 reads; the
interpreter runs them sequentially, i.e. the vectors hold what a JVM with one worker thread computes.  Library containers are shims that follow
the published layouts (jars absent): fastutil Int2ObjectOpenHashMap / IntOpenHashSet (incl. the iterator-driven removeAll), java.util.HashSet,
ConcurrentHashMap (bins in insertion order, a transfer reverses the nodes before a bin's last run: oracle/pyref.chm_key_order)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import minijvm as J  # noqa: E402
from oracle import pyref  # noqa: E402
from oracle import make_ref_hier as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_myclust.npz")


class FuShim:
    def __init__(self):
        self.s = pyref.FuIntSet()


class MVM(H.HVM):
    def collect(self, items, col):
        r = super().collect(items, col)
        if col is not None and col.name == "collector:groupingBy" and isinstance(r.v, dict):      # ConcurrentHashMap iteration order
            order, _ = pyref.chm_key_order([int(k) for k in r.v.keys()])
            r.v = {k: r.v[k] for k in order}
        return r

    def native(self, cls, name, desc, args):
        a = args
        recv = a[0] if a else None
        store = recv.native if isinstance(recv, J.JObj) else (recv.v if isinstance(recv, J.JNative) else None)
        if name == "<init>" and cls.endswith("fastutil/ints/IntOpenHashSet"):
            recv.native = FuShim()
            return None
        if isinstance(store, FuShim):
            s = store.s
            if name == "add":
                return int(s.add(int(a[1])))
            if name == "size":
                return len(s)
            if name == "isEmpty":
                return int(len(s) == 0)
            if name == "contains":
                return int(int(a[1]) in s.order())
            if name == "stream":
                return J.JNative("java/util/stream/Stream", J.JStream(s.order()))
            if name == "removeAll":
                victims = a[1].v if isinstance(a[1], J.JNative) else a[1].native
                before = len(s)
                s.remove_all([int(x) for x in victims])
                return int(len(s) != before)
        return super().native(cls, name, desc, args)


def main():
    n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    H.install_set_extras()
    vm = MVM(H.JARS)
    H.install_overrides(vm)
    rng = np.random.default_rng(20261017)
    t0 = time.time()
    jobs = []
    for t in range(n_jobs):
        n = int(rng.integers(101, 140)) if t % 3 else int(rng.integers(140, 330))
        if t < 4:
            n = int(rng.integers(20, 60))                       # the class itself has no size limit: small jobs run the same code sequentially
        packed = H.random_packed(rng, n, t % 3) if t % 5 else umi_like(rng, n)
        qv = int(rng.integers(0, 2))
        prm = (2, 1, 3000, int(rng.choice([50, 50, 3])))
        res, n_found = H.run_job(vm, packed.tolist(), prm, qv, cls="ClusterOne_MyClustering")
        jobs.append(dict(n=n, packed=packed, qv=qv, prm=prm, res=res, n_found=n_found))
        exp = pyref.assign_myclust(packed.tolist(), prm[0], prm[3], bool(qv))                     # the restatement, for immediate feedback
        bad = [i for i in range(n) if (exp[i]["assigned"], exp[i]["u1"] if exp[i]["assigned"] else -1, exp[i]["u2"] if exp[i]["assigned"] else -1) !=
               (bool(res[i]["assigned"]), res[i]["u1"], res[i]["u2"]) or (exp[i]["assigned"] and "UMI(%d,%d)" % (exp[i]["center"], exp[i]["off_mean"]) != res[i]["u8"])]
        if bad:
            print("    pyref differs on reads", bad[:10], [(exp[i], res[i]) for i in bad[:2]], "tie_unpin", exp[0]["tie_unpin"])
        print("  job %d / %d (n = %d): %d assigned, %.0f s, %d bytecodes" % (t, n_jobs, n, sum(r["assigned"] for r in res), time.time() - t0, vm.n_insn), flush=True)
    off = np.cumsum([0] + [j["n"] for j in jobs]).astype(np.int64)
    moff = np.cumsum([0] + [j["n"] ** 2 for j in jobs]).astype(np.int64)
    flat = lambda k, dt: np.array([r[k] for j in jobs for r in j["res"]], dtype=dt)
    np.savez_compressed(OUT, job_offsets=off, out_offsets=moff, packed=np.concatenate([j["packed"].ravel() for j in jobs]),
                        qv01=np.array([j["qv"] for j in jobs], dtype=np.uint8), params=np.array([j["prm"] for j in jobs], dtype=np.int32),
                        assigned=flat("assigned", np.int8), u8=np.array([r["u8"] for j in jobs for r in j["res"]]), u1=flat("u1", np.int8),
                        u2=flat("u2", np.int8), pos2=flat("pos2", np.int8), flagval=flat("flagval", np.int64),
                        n_found=np.array([j["n_found"] for j in jobs], dtype=np.int32))
    print("ClusterOne_MyClustering.call: %d jobs, %d reads, %d assigned, %.0f s, %d bytecodes" %
          (len(jobs), int(off[-1]), int(flat("assigned", np.int8).sum()), time.time() - t0, vm.n_insn))


def umi_like(rng, n):
    """a matrix shaped like real deep jobs: a few molecules with many reads each (identical or 1-2 errors apart), some chimeric bridges"""
    k = int(rng.integers(3, 12))
    lab = rng.integers(0, k, n)
    err = rng.integers(0, 3, n)                                  # errors of each read against its molecule
    e = np.where(lab[:, None] == lab[None, :], np.minimum(err[:, None] + err[None, :], 5), np.minimum(3 + rng.integers(0, 3, (n, n)), 5))
    bridge = rng.random((n, n)) < 0.01
    e = np.where(bridge, rng.integers(1, 3, (n, n)), e)
    e = np.triu(e, 1)
    e = e + e.T
    p1, p2 = rng.integers(0, 3, (n, n)), rng.integers(0, 3, (n, n))
    up = e | (0x08000000 << p1) | (0x01000000 << p2)
    lo = e | (0x08000000 << p2.T) | (0x01000000 << p1.T)
    packed = np.where(np.arange(n)[:, None] <= np.arange(n)[None, :], up, lo)
    np.fill_diagonal(packed, 0x10000000 | 0x02000000)
    return packed.astype(np.int32)


if __name__ == "__main__":
    main()
