"""Golden vectors of the read grouper FROM THE REFERENCE'S OWN CLASS FILES: ReadGrouper.groupSams
(F!com/rw/umifinder/bamreaders/ReadGrouper.class, ReadGrouper.java:L82-L260) run by oracle/minijvm.py together with ReadGrouper$Cluster,
$ClusterList, $NanoporeReadWithOrderedPosition and BamReader$NanoporeReadChunk as bytecode.  Frozen in tests/golden/ref_grouper.npz.

    python oracle/make_ref_grouper.py [n_cases]

Injected (classes that need htsjdk / the XML binding): a NanoporeRead is a bare field holder whose `sam` answers getFlags() and whose
ReadScanData carries positionOnGenomeForClustering; the BlockingQueue is a list; log4j / Runtime / ManagementFactory calls are dropped.
Library shims used beyond minijvm's own: Arrays.parallelSort(Comparable[]) (stable merge sort by the element's own compareTo bytecode),
IntStream.range, Stream.flatMap / toArray, List.remove(Object) / removeAll (identity equals: the element class defines none),
Math.round(float), Integer.compare / compareTo.  Every case runs on the SAME static region counter, and `chains` of cases replay
BamReader.run's carry-over (BamReader.java:L134-L135: the returned chunk is the head of the next one), so the statefulness is pinned too."""
import functools
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

OUT = os.path.join(ROOT, "tests", "golden", "ref_grouper.npz")
RG = "com/rw/umifinder/bamreaders/ReadGrouper"
CHUNK = "com/rw/umifinder/bamreaders/BamReader$NanoporeReadChunk"
GOPT = "com/google/common/base/Optional"


class FakeClass:
    """a class outside the jars of which the path only reads static fields"""
    def __init__(self, name, statics):
        self.name, self.statics, self.initialized, self.super_name, self.methods = name, statics, True, None, {}


class GVM(H.HVM):
    def __init__(self, jars):
        super().__init__(jars)
        self.fake = {"org/apache/logging/log4j/Level": FakeClass("org/apache/logging/log4j/Level", {k: J.JNative("level:" + k) for k in ("DEBUG", "TRACE", "ALL")})}

    def load(self, name):
        return self.fake.get(name) or super().load(name)

    def new_container(self, supplier):
        if supplier.name == "lambda" and supplier.v[1] != "<init>":          # a Supplier lambda (() -> this, () -> new NanoporeReadChunk(n))
            return self.call_lambda(supplier, [])
        return super().new_container(supplier)

    def collect(self, items, col):
        if col is not None and col.name == "collector:toCollection" and col.v[0].name == "lambda" and col.v[0].v[1] != "<init>":
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
        if isinstance(recv, N) and recv.name == "sam" and name == "getFlags":
            return recv.v
        if isinstance(recv, N) and recv.name == "queue":
            if name == "size":
                return len(recv.v)
            if name == "put":
                recv.v.append(a[1])
                return None
        if cls == "java/lang/Runtime":
            return N("runtime") if name == "getRuntime" else J.L(0)
        if cls == "java/lang/Math" and name == "round" and desc == "(F)I":      # floor(a + 1/2) on the float value, exact
            f = float(a[0])
            return max(-(1 << 31), min((1 << 31) - 1, math.floor(f + 0.5)))
        if cls == "java/lang/Integer" and name == "compare":
            return (a[0] > a[1]) - (a[0] < a[1])
        if cls == "java/lang/Integer" and name == "compareTo":
            if a[1] is None:
                raise J.JavaThrow("java/lang/NullPointerException", "Integer.compareTo")
            return (a[0] > a[1]) - (a[0] < a[1])
        if cls == "java/util/Arrays" and name == "parallelSort":
            a[0].a.sort(key=functools.cmp_to_key(lambda x, y: self.j_compare(x, y)))      # list.sort is stable, like the JDK's object merge sort
            return None
        if cls == "java/util/stream/IntStream" and name == "range":
            return N("java/util/stream/Stream", J.JStream(list(range(a[0], a[1]))))
        if cls in ("java/util/stream/Stream", "java/util/stream/IntStream") and isinstance(store, J.JStream):
            if name == "toArray":
                arr = J.JArr("L", 0, None)
                arr.a = store.run(self)
                return arr
            if name == "flatMap":
                out = []
                for x in store.run(self):
                    out.extend(self.call_functional(a[1], [x]).v.run(self))
                return N("java/util/stream/Stream", J.JStream(out))
        if cls in ("java/util/List", "java/util/ArrayList", "java/util/Collection", "java/util/LinkedList", "java/util/Collections") and isinstance(store, list):
            if name == "remove" and desc == "(Ljava/lang/Object;)Z":
                for i, e in enumerate(store):
                    if self.j_equals(e, a[1]):
                        del store[i]
                        return 1
                return 0
            if name == "removeAll":
                victims = a[1].v if isinstance(a[1], N) else a[1].native
                n0 = len(store)
                store[:] = [e for e in store if not any(self.j_equals(e, v) for v in victims)]
                return int(len(store) != n0)
        return super().native(cls, name, desc, args)


def make_vm():
    H.install_set_extras()
    vm = GVM(H.JARS)
    st = vm.load("com/rw/umifinder/scanstats/ScanStats")                      # only DEC_FORMATTER is read (a debug message)
    st.initialized = True
    st.statics["DEC_FORMATTER"] = J.JNative("java/text/DecimalFormat", ("###,###,###,###",))
    return vm


def make_read(vm, pos, flag, region):
    nr = H.bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead")
    sd = H.bare(vm, "com/rw/umifinder/reads/nanopore/NanoporeRead$ReadScanData")
    sd.f["positionOnGenomeForClustering"] = J.JNative(GOPT, () if pos is None else (int(pos),))
    nr.f["readScanData"] = J.JNative(GOPT, (sd,))
    nr.f["sam"] = J.JNative("sam", int(flag))
    nr.f["genomicRegionNmber"] = J.JNative(GOPT, () if region < 0 else (J.L(int(region)),))
    return nr


def region_counter(vm):
    c = vm.load(RG + "$Cluster")
    vm.init_class(c)
    return int(c.statics["CURRENT_GENOMIC_REGION_ID"])


def run_chunk(vm, reads, keep, max_dist):
    """reads: list of NanoporeRead objects (kept across the calls of a chain).  Returns (thrown, n_done, carried reads, region per read)"""
    vm.call_static(RG, "setMaxGenomeDistance", "(I)V", int(max_dist))
    chunk = vm.construct(CHUNK, "(I)V", 0)
    for r in reads:
        vm.invoke_virtual(CHUNK, "add", "(Ljava/lang/Object;)Z", [chunk, r])
    q = J.JNative("queue", [])
    g = vm.construct(RG, "()V")
    thrown = ""
    try:
        new = vm.call_virtual(g, "groupSams", "(L%s;Ljava/util/concurrent/BlockingQueue;Z)L%s;" % (CHUNK, CHUNK), chunk, q, int(keep))
    except J.JavaThrow as e:
        thrown, new = e.cls, None
    done = q.v[0].native if q.v else []
    carried = list(new.native) if new is not None else []
    if not thrown and reads:
        assert len(q.v) == 1 and all(x is y for x, y in zip(done, reads)) and all(x is y for x, y in zip(carried, reads[len(done):]))
        assert not keep and not carried or len(done) + len(carried) == len(reads)
    reg = [int(r.f["genomicRegionNmber"].v[0]) if r.f["genomicRegionNmber"].v else -1 for r in reads]
    return thrown, len(done), carried, reg


def gen_chunk(rng, kind, n, max_dist):
    """positions (sorted like a coordinate-sorted BAM, then jittered: the clustering position is an end of the alignment, not its start),
    SAM flags, has-position mask"""
    base = int(rng.choice([1_000, 3_000_000, 40_000_000, 180_000_000]))        # above 2^24 the float cast of the centre quantises
    if kind == 0:                                                             # loci of very different depth, gaps around 1 / 2 / 3 x max_dist
        pos, p = [], base
        while len(pos) < n:
            k = int(rng.choice([1, 2, 3, 5, 12, 40]))
            spread = int(rng.choice([max_dist // 5, max_dist, 2 * max_dist, 5 * max_dist]))
            pos += [p + int(x) for x in rng.integers(0, max(1, spread), k)]
            p += spread + int(rng.choice([max_dist - 1, max_dist, max_dist + 1, 2 * max_dist, 3 * max_dist + 7, 20 * max_dist]))
        pos = np.array(pos[:n])
    elif kind == 1:                                                           # one long smear (transcript ends drifting): off-centre removal + merging
        pos = base + np.cumsum(rng.integers(0, max(2, max_dist // 3), n))
    elif kind == 2:                                                           # uniform noise
        pos = base + rng.integers(0, max(1, n * max_dist // 4), n)
    else:                                                                     # two dense loci just under / over 2 x max_dist apart + stragglers
        d = int(rng.choice([max_dist, 2 * max_dist - 1, 2 * max_dist, 2 * max_dist + 1, 3 * max_dist]))
        pos = np.concatenate([base + rng.integers(0, max_dist // 2 + 1, n // 2), base + d + rng.integers(0, max_dist // 2 + 1, n - n // 2 - n // 8),
                              base + rng.integers(-3 * max_dist, 6 * max_dist, n // 8)])
    pos = np.sort(np.maximum(pos, 1))
    pos = pos + rng.integers(-max_dist // 10, max_dist // 10 + 1, len(pos)) * (rng.random(len(pos)) < 0.2)
    flags = np.where(rng.random(len(pos)) < float(rng.choice([0.0, 0.3, 0.5, 1.0])), 16, 0) | rng.choice([0, 256, 2048], len(pos), p=[0.9, 0.05, 0.05])
    has = rng.random(len(pos)) >= float(rng.choice([0.0, 0.0, 0.05]))
    return np.maximum(pos, 1).astype(np.int64), flags.astype(np.int64), has


# case number -> (max_dist, positions, flags or None = all forward, keepDataEnd)
CONSTRUCTED = {
    # two runs that merge completely: the survivor's centre cache is empty and it is the only cluster -> NullPointerException at L173
    8: (500, [1000] * 4 + [1500] * 4, None, 1),
    # the same chunk without keepDataEnd: no exception, one region of 7 reads (the read that opens the gap joins no run)
    12: (500, [1000] * 4 + [1500] * 4, None, 0),
    # ... and with a reverse-strand cluster beside it: sorted() compares, so every centre is filled and nothing throws
    16: (500, [1000] * 4 + [1500] * 4 + [9000, 9001, 9002], [0] * 8 + [16] * 3, 1),
    # a run of two reads is not closed at the gap: reads 40 000 away join it, the off-centre passes then split it again
    20: (500, [100, 101, 40000, 40001, 40002, 40003, 80000, 80001, 80002], None, 0),
    # positions above 2^24: centre = Math.round((float) mean) is a multiple of 16 here; 3 x max_dist rule of keepDataEnd
    24: (500, [200_000_001 + 3 * i for i in range(9)] + [200_004_001 + i for i in range(5)], None, 1),
    # left / right off-centre removal in the same pass: the right test uses the centre computed BEFORE the left removal
    28: (50, [0, 1, 2] + [100 + i for i in range(10)] + [160, 199, 238, 277], None, 0),
    # equal sizes: the LEFT cluster is merged into the right one (isLeftBigger is strict)
    32: (500, [1000] * 3 + [1500] * 4, None, 0),
    # everything on the reverse strand, supplementary / secondary bits set
    36: (120, [5000 + 7 * i for i in range(30)], [16 | (256 if i % 5 == 0 else 0) | (2048 if i % 7 == 0 else 0) for i in range(30)], 1),
}


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    vm = make_vm()
    rng = np.random.default_rng(20261018)
    from oracle import pyref_group as grouping
    t0 = time.time()
    rows = []                                                                 # one row per groupSams call
    carried = []                                                              # (objects, pos, flags, has, region) of the running chain
    n_bad = 0
    for t in range(n_cases):
        max_dist = int(rng.choice([500, 500, 50, 120]))
        n = int(rng.choice([0, 1, 2, 3, 5, 9, 20, 60, 150, 300]))
        if t < 6:
            n = [0, 1, 2, 3, 4, 7][t]
        chain = t % 4 != 0 and carried                                        # three of four cases continue the previous one's carry-over
        pos, flags, has = gen_chunk(rng, t % 4, n, max_dist)
        forced = CONSTRUCTED.get(t)
        if forced is not None:                                                # hand-made chunks (see CONSTRUCTED)
            max_dist, chain = forced[0], False
            pos = np.array(forced[1], dtype=np.int64)
            flags = np.array(forced[2] if forced[2] is not None else [0] * len(pos), dtype=np.int64)
            has = np.ones(len(pos), dtype=bool)
        if chain:
            objs0, pos0, fl0, has0, reg0 = carried
            pos = np.concatenate([pos0, pos + (int(pos0.max()) if len(pos0) else 0)])      # the stream stays (roughly) coordinate-sorted
            flags, has = np.concatenate([fl0, flags]), np.concatenate([has0, has])
        else:
            objs0, reg0 = [], np.zeros(0, dtype=np.int64)
        reg_in = np.concatenate([reg0, np.full(len(pos) - len(reg0), -1, dtype=np.int64)])
        objs = list(objs0) + [make_read(vm, int(pos[i]) if has[i] else None, int(flags[i]), -1) for i in range(len(objs0), len(pos))]
        keep = bool(rng.random() < 0.6) if forced is None else bool(forced[3])
        id0 = region_counter(vm)
        thrown, n_done, car, reg = run_chunk(vm, objs, keep, max_dist)
        id1 = region_counter(vm)
        rows.append(dict(pos=pos, flags=flags, has=has, reg_in=reg_in, keep=keep, max_dist=max_dist, id0=id0, id1=id1, thrown=thrown,
                         n_done=n_done, n_carried=len(car), reg=np.array(reg, dtype=np.int64)))
        # the Python restatement, for immediate feedback
        G = grouping.ReadGrouper(max_dist, id0)
        r2 = reg_in.copy()
        try:
            li = G.group_sams(pos, flags, r2, keep, has)
            exp = ("", 0 if li is None else li + 1, G.next_region_id, r2.tolist())
        except grouping.NullCenterError:
            exp = ("java/lang/NullPointerException", 0, G.next_region_id, None)
        got = (thrown, n_done, id1, reg)
        if exp[:3] != got[:3] or (not thrown and exp[3] != got[3]):
            n_bad += 1
            print("    PYREF DIFFERS case %d: exp %s got %s" % (t, exp[:3], got[:3]))
        k = len(objs) - len(car)
        carried = (car, pos[k:], flags[k:], has[k:], np.array(reg[k:], dtype=np.int64)) if car and not thrown else []
        print("  case %d / %d (n = %d, max %d, keep %d): done %d carried %d regions %d..%d %s, %.0f s, %d bytecodes" %
              (t, n_cases, len(pos), max_dist, keep, n_done, len(car), id0, id1, thrown, time.time() - t0, vm.n_insn), flush=True)
    off = np.cumsum([0] + [len(r["pos"]) for r in rows]).astype(np.int64)
    cat = lambda k, dt: np.concatenate([np.asarray(r[k], dtype=dt) for r in rows]) if rows else np.zeros(0, dtype=dt)
    col = lambda k, dt: np.array([r[k] for r in rows], dtype=dt)
    np.savez_compressed(OUT, offsets=off, position=cat("pos", np.int64), flags=cat("flags", np.int32), has_position=cat("has", np.uint8),
                        region_in=cat("reg_in", np.int64), region_out=cat("reg", np.int64), keep_data_end=col("keep", np.uint8),
                        max_dist=col("max_dist", np.int32), id_before=col("id0", np.int64), id_after=col("id1", np.int64),
                        thrown=np.array([r["thrown"] for r in rows]), n_done=col("n_done", np.int64), n_carried=col("n_carried", np.int64))
    print("ReadGrouper.groupSams: %d calls, %d reads, %d region numbers consumed, %d thrown, pyref differs on %d, %.0f s, %d bytecodes" %
          (len(rows), int(off[-1]), region_counter(vm), sum(bool(r["thrown"]) for r in rows), n_bad, time.time() - t0, vm.n_insn))


if __name__ == "__main__":
    main()
