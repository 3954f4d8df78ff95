"""Golden vectors of the small-job clusterer FROM THE REFERENCE'S OWN CLASS FILES: ClusterOneHierarchical.call
(F!com/rw/umifinder/analyzers/clustering/ClusterOneHierarchical.class) run by oracle/minijvm.py, with LingPipe's CompleteLinkClusterer /
SingleLinkClusterer / Dendrogram / LinkDendrogram / LeafDendrogram / BoundedPriorityQueue (+ Entry, EntryComparator, QueueIterator) / ObjectToSet /
ScoredObject comparators (Jar/lib/Aliasi_ClusteringLib-1.0.jar), DistanceMatrix, OneUmiCluster, ClusterOneBase.setSamflagsAndStatsForClustered,
BestEditDistance, PlusMinusOneEnum and commons-lang3's ImmutablePair as bytecode.  Frozen in tests/golden/ref_hier.npz.

    python oracle/make_ref_hier.py [n_jobs]

Injected (not executed): the packed matrix itself (DistanceMatrix.<init> would call generateDistanceMatrix on OneNanoporeResult objects;
calcEditDistances is pinned separately, ref_umi_pairs.npz) and the SAM side of OneNanoporeResult (setAttribute, the read's strings, the
statistics objects): the driver records the (tag, value) pairs the bytecode writes.  JDK containers are shims: TreeSet (sorted by the
reference's EntryComparator bytecode), HashMap, LinkedList, java.util.HashSet (JDK HashMap iteration order; for elements WITHOUT hashCode() —
LingPipe's PairScore — the JVM orders by identity hash, i.e. arbitrarily: the shim iterates those in insertion order, the canonical order of
oracle and kernel), fastutil IntOpenHashSet (OneUmiCluster's superclass, jar absent: published 8.2.2 layout, oracle/pyref.fastutil_key_order).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import minijvm as J  # noqa: E402
from oracle import pyref  # noqa: E402

REF = "/root/reference/Jar"
JARS = [REF + "/NanoporeBC_UMI_finder-2.1.jar", REF + "/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar", REF + "/lib/Aliasi_ClusteringLib-1.0.jar",
        REF + "/lib/commons-lang3-3.17.0.jar"]
OUT = os.path.join(ROOT, "tests", "golden", "ref_hier.npz")
CL = "com/rw/umifinder/analyzers/clustering/"
ONR = "com/rw/umifinder/reads/nanopore/OneNanoporeResult"


class TreeSetShim:
    """java.util.TreeSet(Comparator): elements kept sorted by the comparator's own bytecode; compare == 0 means 'same element'"""

    def __init__(self, vm, comparator):
        self.vm, self.cmp, self.items = vm, comparator, []

    def _c(self, x, y):
        return self.vm.invoke_virtual(self.cmp.cls.name, "compare", "(Ljava/lang/Object;Ljava/lang/Object;)I", [self.cmp, x, y])

    def _find(self, e):
        lo, hi = 0, len(self.items)
        while lo < hi:                                        # the comparator is a total order on the entries (score, then entry id)
            mid = (lo + hi) // 2
            c = self._c(e, self.items[mid])
            if c == 0:
                return mid, True
            if c < 0:
                hi = mid
            else:
                lo = mid + 1
        return lo, False

    def add(self, e):
        i, found = self._find(e)
        if found:
            return 0
        self.items.insert(i, e)
        return 1

    def remove(self, e):
        i, found = self._find(e)
        if found:
            del self.items[i]
            return 1
        return 0


class IntSetShim:
    """it.unimi.dsi.fastutil.ints.IntOpenHashSet as OneUmiCluster uses it: add / remove / contains / size / iteration order"""

    def __init__(self):
        self.inserted = []

    def order(self):
        return pyref.fastutil_intset_order(self.inserted)


class HVM(J.VM):
    def __init__(self, jars):
        super().__init__(jars)
        self.attrs = {}                                       # read index -> {tag: value} written through OneNanoporeResult.setAttribute
        self.flagged = set()

    def j_hash(self, a):
        if isinstance(a, J.JObj) and "mMember" in a.f:          # SmallSet$SingletonSet: AbstractSet.hashCode = sum of the elements' hash codes
            return self.j_hash(a.f["mMember"])
        if isinstance(a, J.JNative) and isinstance(a.v, J.JdkHashSet):
            return J.i32(sum(self.j_hash(e) for e in a.v.items()))
        return super().j_hash(a)

    def new_container(self, supplier):
        target = supplier.v[0]
        if self.load(target) is not None:                     # a class of the jars (OneUmiCluster::new)
            return self.construct(target, "()V")
        return super().new_container(supplier)

    def collect(self, items, col):
        if col is not None and col.name == "collector:toCollection" and self.load(col.v[0].v[0]) is not None:
            c = self.new_container(col.v[0])
            for x in items:
                self.invoke_virtual(c.cls.name, "add", "(Ljava/lang/Object;)Z", [c, x])
            return c
        return super().collect(items, col)

    def native(self, cls, name, desc, args):
        a = args
        recv = a[0] if a else None
        store = recv.native if isinstance(recv, J.JObj) else (recv.v if isinstance(recv, J.JNative) else None)
        N = J.JNative
        if name == "<init>":
            if cls == "java/util/TreeSet":
                recv.v = TreeSetShim(self, a[1])
                return None
            if cls in ("java/util/HashMap",) and isinstance(recv, J.JObj):
                recv.native = {}
                return None
            if cls == "java/util/HashSet" and len(a) == 2 and isinstance(a[1], int):
                hs = J.JdkHashSet(self)
                if isinstance(recv, J.JNative):
                    recv.v = hs
                else:
                    recv.native = hs
                return None
            if cls.endswith("fastutil/ints/IntOpenHashSet"):
                recv.native = IntSetShim()
                return None
            if cls == "java/util/AbstractSet":
                return None
        if isinstance(store, TreeSetShim):
            if name == "add":
                return store.add(a[1])
            if name == "remove":
                return store.remove(a[1])
            if name in ("first", "last"):
                if not store.items:
                    raise J.JavaThrow("java/util/NoSuchElementException")
                return store.items[0 if name == "first" else -1]
            if name == "isEmpty":
                return int(not store.items)
            if name == "size":
                return len(store.items)
            if name == "iterator":
                return N("java/util/Iterator", [list(store.items), 0, store])
            if name == "clear":
                store.items = []
                return None
        if isinstance(store, IntSetShim):
            if name == "add":
                v = int(a[1])
                if v in store.inserted:
                    return 0
                store.inserted.append(v)
                return 1
            if name == "size":
                return len(store.inserted)
            if name == "isEmpty":
                return int(not store.inserted)
            if name == "contains":
                return int(int(a[1]) in store.inserted)
            if name == "stream":
                return N("java/util/stream/Stream", J.JStream(store.order()))
            if name in ("remove", "rem"):
                v = int(a[1])
                if v in store.inserted:
                    # removal shifts slots (fastutil shiftKeys); the order after a removal is not needed on the <= 100 path
                    raise NotImplementedError("IntOpenHashSet.remove")
                return 0
        if cls == "java/util/Iterator" and isinstance(recv, N) and recv.name == "java/util/Iterator":
            if name == "remove" and len(recv.v) > 2:
                cur = recv.v[0][recv.v[1] - 1]
                ts = recv.v[2]
                if isinstance(ts, TreeSetShim):
                    ts.items = [x for x in ts.items if x is not cur]
                return None
        if isinstance(store, dict) and cls in ("java/util/HashMap", "java/util/Map", "com/aliasi/util/ObjectToSet"):
            if name == "remove":
                return store.pop(a[1], None)
            if name == "keySet":
                return N("java/util/ArrayList", list(store.keys()))
        if isinstance(store, J.JdkHashSet):
            if name == "contains" and hasattr(store, "order"):   # identity-hashed elements: equals() is ==
                return int(any(x is a[1] for x in store.order))
            if name == "remove":
                return store.remove(a[1])
            if name == "iterator":
                return N("java/util/Iterator", [store.items(), 0])
            if name == "toArray":
                arr = a[1] if len(a) > 1 and isinstance(a[1], J.JArr) else J.JArr("L", 0, None)
                arr.a = list(store.items())
                return arr
            if name == "hashCode":                             # AbstractSet.hashCode: sum of the elements' hash codes
                return J.i32(sum(self.j_hash(e) for e in store.items()))
            if name == "equals":
                o = a[1].v if isinstance(a[1], N) else getattr(a[1], "native", None)
                return int(isinstance(o, J.JdkHashSet) and o.size == store.size and all(any(self.j_equals(x, y) for y in o.items()) for x in store.items()))
        if isinstance(store, list) and name == "toArray" and len(a) == 2 and isinstance(a[1], J.JArr):     # toArray(T[]): fills the caller's array
            if len(a[1].a) >= len(store):
                a[1].a[:len(store)] = list(store)
                return a[1]
            arr = J.JArr("L", 0, None)
            arr.a = list(store)
            return arr
        if cls == "java/util/LinkedList" or (isinstance(recv, N) and recv.name == "java/util/LinkedList"):
            if name == "addFirst":
                store.insert(0, a[1])
                return None
            if name == "removeFirst":
                return store.pop(0)
        if cls == "java/util/stream/IntStream" and name == "range":
            return N("java/util/stream/Stream", J.JStream(list(range(a[0], a[1]))))
        if cls in ("java/util/stream/Stream", "java/util/stream/IntStream", "java/util/stream/LongStream") and isinstance(getattr(recv, "v", None), J.JStream):
            if name == "mapToLong":
                return N("java/util/stream/Stream", J.JStream(recv.v.src, recv.v.ops + [("map", a[1])]))
            if name == "findAny":
                r = recv.v.run(self, limit=1)
                return N("java/util/Optional", (r[0],) if r else ())
            if name == "reduce" and len(a) == 3:
                acc = a[1]
                for x in recv.v.run(self):
                    acc = self.call_functional(a[2], [acc, x])
                return acc
            if name == "min" and len(a) == 1:
                v_ = [int(x) for x in recv.v.run(self)]
                return N("java/util/OptionalLong", (J.L(min(v_)),) if v_ else ())
            if name == "flatMap":
                out = []
                for x in recv.v.run(self):
                    s2 = self.call_functional(a[1], [x])
                    out += s2.v.run(self)
                return N("java/util/stream/Stream", J.JStream(out))
        if cls == "java/util/OptionalLong":
            if name == "isPresent":
                return int(len(recv.v) == 1)
            if name == "getAsLong":
                return recv.v[0]
        if cls == "java/lang/Math":
            if name == "round":
                import math
                return J.L(int(math.floor(float(a[0]) + 0.5)))
            if name == "pow":
                return J.D(float(a[0]) ** float(a[1]))
        if cls == "java/lang/Double" and name in ("valueOf", "doubleValue"):
            return a[0]
        if cls == "java/lang/Double" and name == "isNaN":
            return int(a[0] != a[0])
        if cls == "java/lang/Float" and name == "floatValue":
            return a[0]
        if cls == "java/lang/Byte" and name in ("valueOf", "byteValue"):
            return a[0]
        if cls == "java/lang/Integer" and name == "sum":
            return J.i32(a[0] + a[1])
        if cls == "java/lang/Integer" and name == "compareTo":
            return (a[0] > a[1]) - (a[0] < a[1])
        if cls == "java/lang/Long" and name in ("valueOf", "longValue"):
            return a[0]
        if cls == "java/lang/String" and name == "valueOf":
            return str(int(a[0]))
        if cls == "java/lang/Boolean" and name in ("valueOf", "booleanValue"):
            return a[0]
        if name in ("accept",) and isinstance(recv, N) and recv.name == "lambda":
            return self.call_functional(recv, a[1:])
        if cls == "java/lang/System" and name == "currentTimeMillis":
            return J.L(0)
        if cls == "java/util/Arrays" and name == "sort" and len(a) == 2:        # TimSort: stable
            import functools
            cmpo = a[1]
            cmpf = (lambda x, y: self.call_functional(cmpo, [x, y])) if isinstance(cmpo, N) else \
                   (lambda x, y: self.invoke_virtual(cmpo.cls.name, "compare", "(Ljava/lang/Object;Ljava/lang/Object;)I", [cmpo, x, y]))
            a[0].a.sort(key=functools.cmp_to_key(cmpf))
            return None
        return super().native(cls, name, desc, args)


def install_set_extras():
    """java.util.HashSet.remove + the insertion-order iteration of identity-hashed elements (idempotent: a second call in one process would
    wrap add() twice and list every identity-hashed element twice)"""
    if getattr(J.JdkHashSet, "_extras_installed", False):
        return

    def remove(self, e):
        if self.table is None:
            return 0
        for chain in self.table:
            for k, (h, x) in enumerate(chain):
                if x is e or self.vm.j_equals(e, x):
                    del chain[k]
                    self.size -= 1
                    if hasattr(self, "order"):
                        self.order = [y for y in self.order if y is not x]
                    return 1
        return 0

    orig_add, orig_items = J.JdkHashSet.add, J.JdkHashSet.items

    def add(self, e):
        # identity-hashed = a class of the jars whose superclass chain reaches java.lang.Object without defining hashCode() (LingPipe's
        # PairScore); a SmallSet is an AbstractSet and hashes by content
        ident = isinstance(e, J.JObj) and self.vm.find_method(e.cls, "hashCode()I") == ("java/lang/Object", None)
        r = orig_add(self, e)
        if ident:
            if not hasattr(self, "order"):
                self.order = []
            if r:
                self.order.append(e)
        return r

    def items(self):
        if hasattr(self, "order"):
            return list(self.order)
        return orig_items(self)
    J.JdkHashSet.remove, J.JdkHashSet.add, J.JdkHashSet.items = remove, add, items
    J.JdkHashSet._extras_installed = True


def bare(vm, name):
    return vm.new_object(vm.load(name), init=False)


def make_params(vm, ed_complete, ed_single, single_thr, fold):
    P = bare(vm, "com/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams")
    U = bare(vm, "com/rw/parameters/UMIparameters")
    U.f.update(umi_length=12, umi_completelinkclusteringED=ed_complete, umi_singlelinkclusteringED=ed_single,
               complexity_threshold_for_switch_to_single_link_clustering=single_thr, foldDepthBelowMaxDiscardForClustering=fold)
    P.f["umis"] = U
    tags = bare(vm, "com/rw/umifinder/flags/OutputSAMtags")
    ut = bare(vm, "com/rw/umifinder/flags/OutputSAMtags$UmiFindingSamTags")
    for fld, tag in (("UMI_SEQ", "U8"), ("UMI_READSEQ", "U7"), ("UMI_IS_FROM_CLUSTERING", "UC"), ("UMI_ED", "U1"), ("UMI_ED_SECOND_BEST_MATCH", "U2")):
        fl = bare(vm, "com/rw/umifinder/flags/OutputSAMtags$OneSamFlag")
        fl.f["samFlag"] = tag
        ut.f[fld] = fl
    tags.f["umiFindingSamTags"] = ut
    P.f["samFlags"] = tags
    return P


def run_job(vm, packed, prm, qv01, cls="ClusterOneHierarchical"):
    n = len(packed)
    P = make_params(vm, *prm)
    reads = []
    for i in range(n):
        r = bare(vm, ONR)
        r.f["userObject"] = J.JNative("java/util/Optional", ())
        r.f["umi"] = None
        r.f["umiFindingFlagValue"] = 0
        r.f["$idx"] = i
        nr = bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead")
        sd = bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead$ReadScanData")
        sd.f["mean_qv"] = (20.0 if qv01 else 10.0) if i == 0 else 15.0       # mean_qv(read 0) > mean_qv(read 1) iff qv01
        nr.f["readScanData"] = J.JNative("com/google/common/base/Optional", (sd,))
        r.f["nanoporeRead"] = nr
        reads.append(r)
    pair = vm.construct("org/apache/commons/lang3/tuple/ImmutablePair", "(Ljava/lang/Object;Ljava/lang/Object;)V", 0, J.JNative("java/util/ArrayList", reads))
    stats = bare(vm, "com/rw/umifinder/scanstats/ScanStats")
    stats.f["nUMIfoundClustering"] = J.JNative("java/util/concurrent/atomic/AtomicInteger", [0])
    vm.attrs, vm.flagged, vm.packed = {}, set(), packed
    me = vm.construct(CL + cls, "(Lcom/rw/umifinder/parameters/ParametersBarcodeUMiFinderAppParams;Lorg/apache/commons/lang3/tuple/ImmutablePair;"
                      "Lcom/rw/umifinder/scanstats/ScanStats;)V", P, pair, stats)
    vm.call_virtual(me, "call", "()Lorg/apache/commons/lang3/tuple/ImmutablePair;")
    skipped_flag = None
    out = []
    for i, r in enumerate(reads):
        at = vm.attrs.get(i, {})
        out.append(dict(assigned=int("U1" in at), u8=at.get("U8", ""), u1=int(at.get("U1", -1)), u2=int(at.get("U2", -1)) if "U2" in at else -1,
                        pos2=at.get("pos2", -1), flagval=int(r.f["umiFindingFlagValue"])))
    return out, int(stats.f["nUMIfoundClustering"].v[0])


def stub_stats_enum(vm):
    """OneReadOrSamScanStats$Flags: its static initialiser builds EnumSets of every statistics flag (reporting, out of scope); the path only
    passes four constants around and reads one mask"""
    c = vm.load("com/rw/umifinder/scanstats/OneReadOrSamScanStats$Flags")
    c.initialized = True
    for k in ("UMI_FOUND_IN_CLUSTERING", "UMI_TOT_FOUND", "UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_MINUSONE", "UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_PLUSONE",
              "UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_ZERO"):
        o = J.JObj(c)
        o.f["$name"] = k
        c.statics[k] = o
    c.statics["UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_SET"] = J.L(7 << 40)


def install_overrides(vm):
    """the SAM / statistics side of OneNanoporeResult and the matrix injection"""
    stub_stats_enum(vm)
    real_run = vm.run

    def run(c, key, args):
        nm = key.split("(")[0]
        if c.name == "com/rw/clustering/DistanceMatrix" and nm == "<init>":
            me, dat, umi_len, params, gen = args
            packed = vm.packed
            n = len(packed)
            rows = J.JArr("[", n, None)
            for a_ in range(n):
                row = J.JArr("L", n, None)
                for b_ in range(n):
                    cell = bare(vm, "com/rw/clustering/ClusteringEditDistanceBase")
                    best = bare(vm, "com/rw/clustering/ClusteringEditDistanceBase$BestEditDistance")
                    best.f["ed"] = int(packed[a_][b_])
                    cell.f["bestEditDistance"] = best
                    row.a[b_] = cell
                rows.a[a_] = row
            me.f["nanoporeData"] = dat.f["right"]
            me.f["wasPregrouped"] = 0
            me.f["parameters"] = params
            me.f["distanceMatrix"] = rows
            iwn = real_run(c, "generateIndicesWithNeighbours()Ljava/util/List;", [me]) if gen else None
            me.f["indicesWithNeighbors"] = J.JNative("java/util/Optional", (iwn,) if gen else ())
            me.f["indicesPregrouped"] = J.JNative("java/util/Optional", ())
            return None
        if c.name == ONR:
            if nm == "setAttribute":
                vm.attrs.setdefault(args[0].f["$idx"], {})[args[1]] = args[2]
                return None
            if nm == "getPostBCUMIseqOffset":
                return J.JNative("java/util/Optional", (J.JNative("umi", "UMI(%d,%d)" % (args[0].f["$idx"], args[2])),))
            if nm == "getPostBCUMIseq":
                return J.JNative("java/util/Optional", (J.JNative("umi", "READ(%d)" % args[0].f["$idx"]),))
            if nm == "setReadAndSamStatFlag":
                return None
            if nm == "getOneReadScanStats":
                return J.JNative("stats", args[0].f["$idx"])
        return real_run(c, key, args)
    vm.run = run
    real_native = vm.native

    def native(cls, name, desc, args):
        a = args
        if a and isinstance(a[0], J.JNative) and a[0].name == "umi" and name == "toString":
            return a[0].v
        if a and isinstance(a[0], J.JNative) and a[0].name == "stats":
            if name == "getFlag":
                return J.L(0)
            if name == "setFlag":                            # UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_{MINUSONE, PLUSONE, ZERO}
                vm.attrs.setdefault(a[0].v, {})["pos2"] = {"UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_MINUSONE": 0, "UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_ZERO": 1,
                                                           "UMI_FOUND_IN_CLUSTERING_PREDICTED_POS_PLUSONE": 2}[a[1].f["$name"]]
                return None
            if name == "getClusteringUMIeditDistancesCreateIfNull":
                return J.JNative("eds", a[0].v)
        if a and isinstance(a[0], J.JNative) and a[0].name == "eds" and name == "setEditDistanceIfSmaller":
            return None
        return real_native(cls, name, desc, args)
    vm.native = native


def random_packed(rng, n, mode):
    """as tests/test_umi_assign.random_packed"""
    if mode == 0:
        lab = rng.integers(0, max(1, n // 3), n)
        base = np.where(lab[:, None] == lab[None, :], rng.integers(0, 3, (n, n)), rng.integers(2, 6, (n, n)))
    elif mode == 1:
        lab = rng.integers(0, max(1, n // 4), n)
        base = np.where(lab[:, None] == lab[None, :], rng.integers(0, 2, (n, n)), 5)
    else:
        pc = float(rng.choice([0.05, 0.2, 0.5, 0.9]))
        base = np.where(rng.random((n, n)) < pc, rng.integers(0, 3, (n, n)), rng.integers(3, 6, (n, n)))
    e = np.triu(base, 1)
    e = e + e.T
    p1, p2 = rng.integers(0, 3, (n, n)), rng.integers(0, 3, (n, n))
    up = e | (0x08000000 << p1) | (0x01000000 << p2)
    lo = e | (0x08000000 << p2.T) | (0x01000000 << p1.T)
    packed = np.where(np.arange(n)[:, None] <= np.arange(n)[None, :], up, lo)
    np.fill_diagonal(packed, 0x10000000 | 0x02000000)
    return packed.astype(np.int32)


def main():
    n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 220
    install_set_extras()
    vm = HVM(JARS)
    install_overrides(vm)
    rng = np.random.default_rng(31337)
    t0 = time.time()
    jobs = []
    for t in range(n_jobs):
        n = int(rng.integers(2, 14)) if t % 4 else int(rng.integers(14, 101 if t % 8 == 0 else 40))
        packed = random_packed(rng, n, t % 3)
        qv = int(rng.integers(0, 2))
        prm = (2, 1, int(rng.choice([3000, 3000, 3000, 5])), int(rng.choice([50, 50, 2])))
        res, n_found = run_job(vm, packed.tolist(), prm, qv)
        jobs.append(dict(n=n, packed=packed, qv=qv, prm=prm, res=res, n_found=n_found))
        if t % 20 == 0:
            print("  job %d / %d (n = %d), %.0f s, %d bytecodes" % (t, n_jobs, n, time.time() - t0, vm.n_insn), flush=True)
    off = np.cumsum([0] + [j["n"] for j in jobs]).astype(np.int64)
    moff = np.cumsum([0] + [j["n"] ** 2 for j in jobs]).astype(np.int64)
    flat = lambda k, dt: np.array([r[k] for j in jobs for r in j["res"]], dtype=dt)
    np.savez_compressed(OUT, job_offsets=off, out_offsets=moff, packed=np.concatenate([j["packed"].ravel() for j in jobs]),
                        qv01=np.array([j["qv"] for j in jobs], dtype=np.uint8), params=np.array([j["prm"] for j in jobs], dtype=np.int32),
                        assigned=flat("assigned", np.int8), u8=np.array([r["u8"] for j in jobs for r in j["res"]]), u1=flat("u1", np.int8),
                        u2=flat("u2", np.int8), pos2=flat("pos2", np.int8), flagval=flat("flagval", np.int64),
                        n_found=np.array([j["n_found"] for j in jobs], dtype=np.int32))
    print("ClusterOneHierarchical.call: %d jobs, %d reads, %d assigned, %.0f s, %d bytecodes" %
          (len(jobs), int(off[-1]), int(flat("assigned", np.int8).sum()), time.time() - t0, vm.n_insn))


if __name__ == "__main__":
    main()
