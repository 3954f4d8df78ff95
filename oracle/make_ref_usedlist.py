"""Golden vectors of the hand-over between the two passes of scanfastq FROM THE REFERENCE'S OWN CLASS FILES:
UsedBarcodesListData.filterLowCounts (lambda$new$4 / $0, UsedCellBCListGenerator.java:L359-L363) and
BarcodeDatasetColissionTester.generateColissionMergedBCmap (…java:L158-L203) run by oracle/minijvm.py on Matches objects that the reference's own
BarcodeMatchTester.doJob produced for every barcode of the list (the collision tester's settings, …java:L215-L222).
Frozen in tests/golden/ref_usedlist.npz.

    python oracle/make_ref_usedlist.py [n_cases]

Injected: the tester object is a bare field holder whose colissionsFromScan is already filled (so submitJobs() returns at once: no thread
pool); barcodes_b4filtering is a dict-backed shim of fastutil's Long2ObjectMap (get / containsKey / remove / entry set: only order-free uses);
java.util.HashMap<Long, Set> iterates in the JDK's order as modelled by minijvm.JdkHashSet (bins by index, a bin in insertion order)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import minijvm as J  # noqa: E402
from oracle import make_ref_vectors as M  # noqa: E402
from oracle import make_ref_hier as H  # noqa: E402
from oracle import pyref_usedlist as P  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_usedlist.npz")
CT = "com/rw/nanoporereadscanner/analyzers/BarcodeDatasetColissionTester"
UL = "com/rw/nanoporereadscanner/analyzers/UsedCellBCListGenerator$UsedBarcodesListData"
COLLIDE = np.dtype([("bc", "<u8", (2,)), ("valid", "u1"), ("n_sub", "u1", (2,)), ("n_ins", "u1", (2,)), ("n_del", "u1", (2,)), ("pad", "u1")])


class FuMap:
    """Long2ObjectMap<AtomicInteger> / Long2ObjectOpenHashMap<Integer> stand-in (key order = insertion order; no caller depends on it)"""
    def __init__(self, d=None):
        self.d = dict(d or {})


class OrderedJdkMap(dict):
    """java.util.HashMap<Long, ?>: a dict whose entry set is read in the JDK's iteration order"""


class FakeClass:
    """a class outside the jars of which the path only reads static fields"""
    def __init__(self, name, statics):
        self.name, self.statics, self.initialized, self.super_name, self.methods = name, statics, True, None, {}


class UVM(H.HVM):
    def load(self, name):
        if name == "java/lang/System":
            return FakeClass(name, {"out": J.JNative("logger")})          # System.out.println(progress message): dropped like a log call
        return super().load(name)

    def new_container(self, supplier):
        if supplier.v[0].endswith("Long2ObjectOpenHashMap"):
            return J.JNative(supplier.v[0], FuMap())
        return super().new_container(supplier)

    def collect(self, items, col):
        if col is not None and col.name == "collector:toMap" and col.v[3].v[0].endswith("Long2ObjectOpenHashMap"):
            kf, vf, merge, sup = col.v
            c = self.new_container(sup)
            for x in items:
                k_, v_ = int(self.call_functional(kf, [x])), self.call_functional(vf, [x])
                c.v.d[k_] = self.call_functional(merge, [c.v.d[k_], v_]) if k_ in c.v.d else v_
            return c
        return super().collect(items, col)

    def native(self, cls, name, desc, args):
        a = args
        N = J.JNative
        recv = a[0] if a else None
        store = recv.v if isinstance(recv, N) else None
        if name == "<init>" and cls == "java/util/HashMap" and isinstance(recv, N):
            recv.v = OrderedJdkMap()
            return None
        if isinstance(store, OrderedJdkMap) and name == "entrySet":
            hs = J.JdkHashSet(self)                              # the JDK HashMap algorithm of minijvm, keyed by the Long's own hashCode
            for k in store.keys():
                hs.add(k)
            return N("java/util/ArrayList", [N("entry", (k, store[k])) for k in hs.items()])
        if isinstance(store, FuMap):
            d = store.d
            if name == "get":
                return d.get(int(a[1]))
            if name == "containsKey":
                return int(int(a[1]) in d)
            if name == "remove":
                return d.pop(int(a[1]), None)
            if name == "long2ObjectEntrySet":
                return N("java/util/ArrayList", [N("entry", (J.L(k), v)) for k, v in d.items()])
            if name == "size":
                return len(d)
        if isinstance(recv, N) and recv.name == "entry" and name == "getLongKey":
            return recv.v[0]
        if cls == "java/util/concurrent/atomic/AtomicInteger" and name == "intValue":
            return recv.v[0]
        if cls == "java/util/concurrent/ConcurrentHashMap" and name == "entrySet" and isinstance(store, dict):
            return N("java/util/ArrayList", [N("entry", (k, v)) for k, v in store.items()])
        if cls == "java/lang/Integer" and name == "compare":
            return (a[0] > a[1]) - (a[0] < a[1])
        if cls == "java/lang/Float" and name in ("valueOf", "floatValue"):
            return a[0]
        if cls == "java/lang/Long" and name == "longValue":
            return a[0]
        return super().native(cls, name, desc, args)


def atom(v):
    return J.JNative("java/util/concurrent/atomic/AtomicInteger", [int(v)])


def do_job(vm, keys, w, ed):
    """new BarcodeMatchTester(Optional.of(seq), ed, skipFullMatches, allowIndels, keySet, offset 0, 16, postSeq = null, !doNext).call()'s job: doJob(seq)"""
    t = vm.construct(M.BMT, "(IZZLjava/util/Set;SILcom/rw/nuc/encoding/onebyte/NucleicAcidInmutableOneBytePerBase;Z)V", ed, 1, 1, J.PySet(keys), 0, 16, None, 0)
    seq = vm.construct(M.T2, "(JI)V", J.L(w), 16)
    return vm.call_virtual(t, "doJob", "(Lcom/rw/nuc/encoding/TwoBit/NucleicAcidTwoBitPerBase;)Lcom/rw/nanoporereadscanner/analyzers/BarcodeMatchTester$Matches;", seq)


def run_case(vm, barcodes, counts, ed, factor, fold):
    chm = {}
    collide = np.zeros(len(barcodes), dtype=COLLIDE)
    for i, b in enumerate(barcodes):
        m = do_job(vm, barcodes, int(b), ed)
        items = m.native.items() if m is not None else []
        for o in items:
            e = int(o.f["editDistance"]) - 1
            collide[i]["bc"][e] = int(o.f["matchingBC"]) & M.M64
            collide[i]["valid"] |= 1 << e
            collide[i]["n_sub"][e], collide[i]["n_ins"][e], collide[i]["n_del"][e] = o.f["substitutions"], o.f["insertions"], o.f["deletions"]
        if items:                                              # FutCallBack.onSuccess (…java:L240-L241): only a non-empty Matches is kept
            chm[J.L(int(b))] = m
    # the reference's callbacks fill the ConcurrentHashMap from many threads; its entry set is then sorted by count (stable): present it in input order
    me = H.bare(vm, CT)
    me.f["colissionsFromScan"] = J.JNative("java/util/concurrent/ConcurrentHashMap", chm)
    me.f["barcodes_b4filtering"] = J.JNative("fu", FuMap({int(b): atom(c) for b, c in zip(barcodes, counts)}))
    par = H.bare(vm, "com/rw/nanoporereadscanner/parameters/ParametersReadScannerApp")
    rsp = H.bare(vm, "com/rw/parameters/ReadScannerParameters")
    rsp.f["cellsWithReadsnFoldBelowMaxToKeep"] = int(fold)
    par.f["readScannerParameters"] = rsp
    me.f["params"] = par
    res = vm.call_virtual(me, "generateColissionMergedBCmap", "(II)Lit/unimi/dsi/fastutil/longs/Long2ObjectOpenHashMap;", int(factor), int(ed))
    return collide, {int(k): int(v) for k, v in res.v.d.items()}


def run_filter(vm, counts, record_count):
    """filterLowCounts.apply(map, 2.0f * recordCount / 5000000.0f) as finalizeData calls it (UsedCellBCListGenerator.java:L391-L392)"""
    cutoff = J.np_float32(J.np_float32(2.0) * J.np_float32(int(record_count))) / J.np_float32(5000000.0)
    src = J.JNative("fu", FuMap({i: atom(c) for i, c in enumerate(counts)}))
    out = vm.invoke_exact(UL, "lambda$new$4", "(Lit/unimi/dsi/fastutil/longs/Long2ObjectMap;Ljava/lang/Float;)Lit/unimi/dsi/fastutil/longs/Long2ObjectMap;",
                          [src, J.np_float32(cutoff)])
    keep = np.zeros(len(counts), dtype=bool)
    keep[[int(k) for k in out.v.d.keys()]] = True
    return keep


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    H.install_set_extras()
    vm = UVM(H.JARS)
    c = vm.load(CT)
    c.initialized = True
    c.statics["LOGGER"] = J.JNative("logger")
    rng = np.random.default_rng(20261021)
    t0 = time.time()
    rows, bad = [], 0
    for t in range(n_cases):
        ed = 2 if t % 6 == 5 else 1
        n = int(rng.integers(8, 15)) if ed == 2 else int(rng.choice([5, 20, 40, 90, 140]))
        seeds = [M.rseq(rng, 16) for _ in range(max(2, n // 4))]
        bcs = set()
        while len(bcs) < n:                                    # families of close barcodes: chains of collisions at ED 1 / 2
            s = seeds[int(rng.integers(0, len(seeds)))]
            bcs.add(M.pack((M.mutate(rng, s, int(rng.integers(0, ed + 1))) + M.rseq(rng, 2))[:16]))
        barcodes = np.array(sorted(bcs), dtype=np.uint64)
        rng.shuffle(barcodes)
        counts = np.maximum(1, (10 ** rng.uniform(0, 4.5, n)).astype(np.int64))      # real lists: a few big cells, a long tail
        if t % 4 == 0:
            counts[: n // 3] = counts[0]                       # equal counts: rank ties, stable-sort ties
        factor, fold = int(rng.choice([10, 10, 3, 1])), int(rng.choice([500, 50, 5]))
        collide, final = run_case(vm, [int(b) for b in barcodes], counts, ed, factor, fold)
        rec = int(rng.choice([0, 10 ** 6, 10 ** 7, 10 ** 8]))
        fkeep = run_filter(vm, counts, rec)
        rows.append(dict(barcodes=barcodes, counts=counts.astype(np.int32), collide=collide, ed=ed, factor=factor, fold=fold,
                         final=np.array([int(b) in final for b in barcodes], dtype=np.uint8), record_count=rec, fkeep=fkeep.astype(np.uint8)))
        keep, rank, flags = P.merge_collisions(barcodes, counts, collide, factor, ed, fold)
        if not np.array_equal(keep, rows[-1]["final"].astype(bool)) or any(final[int(b)] != int(c) for b, c, k in zip(barcodes, counts, keep) if k) or \
                not np.array_equal(P.filter_low_counts(counts, rec), fkeep):
            bad += 1
            print("    PYREF DIFFERS case", t)
        print("  case %d / %d: %d barcodes, ED %d, fold %d / %d: %d with matches, %d kept (flags %d); count filter keeps %d; %.0f s" %
              (t, n_cases, n, ed, factor, fold, int((collide["valid"] != 0).sum()), int(keep.sum()), flags, int(fkeep.sum()), time.time() - t0), flush=True)
    off = np.cumsum([0] + [len(r["barcodes"]) for r in rows]).astype(np.int64)
    cat = lambda k: np.concatenate([r[k] for r in rows])
    col = lambda k, dt: np.array([r[k] for r in rows], dtype=dt)
    np.savez_compressed(OUT, offsets=off, barcodes=cat("barcodes"), counts=cat("counts"), collide=cat("collide").view(np.uint8).reshape(-1, COLLIDE.itemsize),
                        ed=col("ed", np.int32), min_count_fold=col("factor", np.int32), cells_fold=col("fold", np.int32), kept=cat("final"),
                        record_count=col("record_count", np.int64), count_filter_keep=cat("fkeep"))
    print("generateColissionMergedBCmap + filterLowCounts: %d cases, %d barcodes, pyref differs on %d, %.0f s, %d bytecodes" %
          (len(rows), int(off[-1]), bad, time.time() - t0, vm.n_insn))


if __name__ == "__main__":
    main()
