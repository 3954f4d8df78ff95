"""Golden vectors of the job former FROM THE REFERENCE'S OWN CLASS FILES: UmiClustering.cluster up to the point where it hands the jobs to its
Submitter thread (F!com/rw/umifinder/analyzers/clustering/UmiClustering.class, UmiClustering.java:L126-L145): groupDataByCellAndRegion
(L97-L118: reads with a cell barcode AND a genomic-region number, grouped by barcode, then by region), the size filter (L135), the split of
oversized groups (L136-L142) — run by oracle/minijvm.py with OneNanoporeResult's static predicates (hasCellBC, getCellBC), ReadScanResult,
BarcodeResult and commons-collections4's ListUtils.partition as bytecode.  Frozen in tests/golden/ref_jobs.npz.

    python oracle/make_ref_jobs.py [n_cases]

Injected: reads are bare field holders (barcode sequence, region number); the Submitter's constructor is intercepted — it receives the job list,
which is what is recorded — and ends the call.  The class's stream is parallel and its maps are ConcurrentHashMaps: the interpreter runs
it sequentially (reads inside a job in input order) and the ORDER OF THE JOBS (map iteration order) is not recorded — no per-read result depends
on it; the vectors hold the set of jobs."""
import glob
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import minijvm as J  # noqa: E402
from oracle import make_ref_hier as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_jobs.npz")
UC = "com/rw/umifinder/analyzers/clustering/UmiClustering"
GOPT = "com/google/common/base/Optional"


class Captured(Exception):
    def __init__(self, jobs):
        self.jobs = jobs


class JVM(H.HVM):
    def native(self, cls, name, desc, args):
        a = args
        N = J.JNative
        recv = a[0] if a else None
        if cls == "java/lang/Math" and name in ("sqrt", "ceil"):
            return J.D(math.sqrt(float(a[0])) if name == "sqrt" else math.ceil(float(a[0])))
        if cls == "java/util/Arrays" and name == "stream":
            return N("java/util/stream/Stream", J.JStream(list(a[0].a)))
        if name == "and" and isinstance(recv, N) and recv.name == "lambda":                      # Predicate.and
            p, q = a[0], a[1]
            return N("pyfunc", lambda x: int(bool(self.call_functional(p, [x])) and bool(self.call_functional(q, [x]))))
        if name == "keySet" and isinstance(recv, N) and isinstance(recv.v, dict):
            return N("java/util/ArrayList", list(recv.v.keys()))
        if name == "stream" and isinstance(recv, J.JObj) and cls == "java/util/AbstractList":     # ListUtils$Partition: AbstractList over get / size
            n_ = self.invoke_virtual(recv.cls.name, "size", "()I", [recv])
            return N("java/util/stream/Stream", J.JStream([self.invoke_virtual(recv.cls.name, "get", "(I)Ljava/lang/Object;", [recv, i]) for i in range(n_)]))
        if name == "subList" and isinstance(recv, N) and isinstance(recv.v, list):
            return N("java/util/ArrayList", recv.v[a[1]:a[2]])
        if cls in ("java/util/stream/Stream",) and isinstance(getattr(recv, "v", None), J.JStream) and name == "flatMap":
            out = []
            for x in recv.v.run(self):
                out.extend(self.call_functional(a[1], [x]).v.run(self))
            return N("java/util/stream/Stream", J.JStream(out))
        return super().native(cls, name, desc, args)

    def j_equals(self, a, b):                                                                 # Long.equals inside distinct()
        if isinstance(a, int) and isinstance(b, int):
            return int(a) == int(b)
        return super().j_equals(a, b)


def make_vm(ram):
    H.install_set_extras()
    jars = H.JARS + glob.glob("/root/reference/Jar/lib/commons-collections4*.jar")
    vm = JVM(jars)
    c = vm.load(UC)
    c.initialized = True                                                                      # <clinit> reads Runtime.maxMemory and builds loggers
    c.statics["LOGGER"] = J.JNative("logger")
    c.statics["MAX_SQUARE_NRECORDSPROCESSING"] = int(int(ram) // 300)                           # RAM_RESERVED / 300 (UmiClustering.java:L59)
    real_run = vm.run

    def run(k, key, args):
        if k.name == UC + "$Submitter" and key.startswith("<init>"):
            raise Captured(args[2])
        return real_run(k, key, args)
    vm.run = run
    return vm


def make_read(vm, idx, bc, region):
    r = H.bare(vm, H.ONR)
    r.f["$idx"] = idx
    nr = H.bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead")
    if bc == -2:                                                                              # no scan data at all
        nr.f["readScanData"] = J.JNative(GOPT, ())
    else:
        sd = H.bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead$ReadScanData")
        if bc == -1:                                                                          # scanned, no barcode found
            sd.f["barcode_Result"] = None
        else:
            br = H.bare(vm, "com/rw/nanoporereadscanner/readerwriter/ReadScanResult$BarcodeResult")
            seq = H.bare(vm, "com/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase")
            seq.f["sequence"] = J.L(int(bc))
            br.f["barcodeseq"] = seq
            sd.f["barcode_Result"] = br
        nr.f["readScanData"] = J.JNative(GOPT, (sd,))
    nr.f["genomicRegionNmber"] = J.JNative(GOPT, () if region < 0 else (J.L(int(region)),))
    r.f["nanoporeRead"] = nr
    return r


def run_case(vm, bc, region):
    arr = J.JArr("L", 0, None)
    arr.a = [make_read(vm, i, int(bc[i]), int(region[i])) for i in range(len(bc))]
    me = H.bare(vm, UC)
    try:
        vm.call_virtual(me, "cluster", "([L%s;)V" % H.ONR, arr)
    except Captured as c:
        jobs = c.jobs.v if isinstance(c.jobs, J.JNative) else c.jobs.native
        out = []
        for j in jobs:
            items = j.v if isinstance(j, J.JNative) else [vm.invoke_virtual(j.cls.name, "get", "(I)Ljava/lang/Object;", [j, i])
                                                          for i in range(vm.invoke_virtual(j.cls.name, "size", "()I", [j]))]
            out.append([x.f["$idx"] for x in items])
        return out
    raise AssertionError("cluster() returned without building its Submitter")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(20261019)
    import __graft_entry__ as g
    pkg = g.load_package()
    t0 = time.time()
    rows = []
    n_bad = 0
    for t in range(n_cases):
        ram = [64e9, 64e9, 8e9, 3e6, 3e5][t % 5]                                              # the small values make the split fire on small groups
        vm = make_vm(ram)
        n = int(rng.choice([0, 1, 2, 10, 80, 400, 1500]))
        n_cells = int(rng.choice([1, 3, 20, 200]))
        n_regions = int(rng.choice([1, 2, 10, 100]))
        cells = rng.integers(0, 1 << 32, n_cells, dtype=np.int64)
        if t % 7 == 0 and n_cells > 1:
            cells[1] = cells[0] + (1 << 32)                                                   # same Long.hashCode low bits, different barcode
        bc = cells[rng.integers(0, n_cells, n)] if n else np.zeros(0, dtype=np.int64)
        region = rng.integers(0, n_regions, n).astype(np.int64) + int(rng.choice([0, 5_000_000_000]))      # region numbers are longs
        miss = rng.random(n)
        bc = np.where(miss < 0.05, -1, np.where(miss < 0.08, -2, bc))
        region = np.where(rng.random(n) < 0.1, -1, region)
        jobs = run_case(vm, bc, region)
        rows.append(dict(ram=int(ram), bc=bc, region=region, jobs=jobs))
        # the mirror, for immediate feedback: group_by_cell_and_region + split_oversized_group
        valid = (bc >= 0) & (region >= 0)
        order, off = pkg.group_by_cell_and_region(bc.astype(np.uint64), region, valid)
        exp = []
        for j in range(len(off) - 1):
            ids = order[off[j]:off[j + 1]].tolist()
            a = 0
            for sz in pkg.split_oversized_group(len(ids), int(ram)):
                exp.append(ids[a:a + sz])
                a += sz
        if sorted(exp) != sorted(jobs):
            n_bad += 1
            print("    MIRROR DIFFERS case %d: %d vs %d jobs" % (t, len(exp), len(jobs)))
        print("  case %d / %d (n = %d, %d cells x %d regions, RAM %.0e): %d jobs, largest %d, %.0f s, %d bytecodes" %
              (t, n_cases, n, n_cells, n_regions, ram, len(jobs), max([len(j) for j in jobs] + [0]), time.time() - t0, vm.n_insn), flush=True)
    off = np.cumsum([0] + [len(r["bc"]) for r in rows]).astype(np.int64)
    joff, jreads, jcase = [0], [], []
    for c, r in enumerate(rows):
        for j in sorted(r["jobs"]):
            jreads += j
            joff.append(len(jreads))
            jcase.append(c)
    cat = lambda k: np.concatenate([np.asarray(r[k], dtype=np.int64) for r in rows])
    np.savez_compressed(OUT, offsets=off, barcode=cat("bc"), region=cat("region"), ram=np.array([r["ram"] for r in rows], dtype=np.int64),
                        job_offsets=np.array(joff, dtype=np.int64), job_reads=np.array(jreads, dtype=np.int64), job_case=np.array(jcase, dtype=np.int64))
    print("UmiClustering.cluster (job former): %d cases, %d reads, %d jobs, mirror differs on %d, %.0f s" %
          (len(rows), int(off[-1]), len(jcase), n_bad, time.time() - t0))


if __name__ == "__main__":
    main()
