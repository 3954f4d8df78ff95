"""A small JVM-bytecode interpreter — TEST INFRASTRUCTURE that lets the REFERENCE'S OWN CLASS FILES run in a container without a JDK.

The hot path of SiCeLoRe 2.1 ships only as bytecode (Jar/NanoporeBC_UMI_finder-2.1.jar, Jar/lib/TwoFourBitNucAcidLibraryMaven-1.0.jar).
This module loads those class files (parser: tools/jdis.py) and interprets the subset of the JVM instruction set they use on that path
(int / long arithmetic with Java's wrap-around and masked shifts, arrays, objects, static initialisers, virtual dispatch).  JDK and
third-party classes that are not in the two jars (java.util.ArrayDeque / ArrayList / HashSet / Optional, boxing, eclipse-collections
IntHashSet / LongHashSet, fastutil LongSet, the Illumina data holders) are modelled by small Python shims that implement exactly the
methods the path calls — membership and ordering semantics only.

It is used by oracle/make_ref_vectors.py (run in the build container, where /root/reference is mounted) to produce golden vectors
FROM THE REFERENCE ITSELF: tests/golden/ref_*.npz.  Nothing at run time (tests on the GPU box, smoke, bench) needs the jars.
"""
import os
import struct
import sys
import zipfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import jdis  # noqa: E402


def np_float32(x):
    return struct.unpack("f", struct.pack("f", float(x)))[0]


class L(int):
    """a Java long on the operand stack / in a local (category-2 value)"""
    __slots__ = ()


class D(float):
    __slots__ = ()


def i32(x):
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def i64(x):
    x &= 0xFFFFFFFFFFFFFFFF
    return L(x - 0x10000000000000000 if x & 0x8000000000000000 else x)


class JavaThrow(Exception):
    def __init__(self, cls, msg=""):
        super().__init__("%s: %s" % (cls, msg))
        self.cls = cls


class JObj:
    __slots__ = ("cls", "f", "native")

    def __init__(self, cls):
        self.cls, self.f, self.native = cls, {}, None

    def __repr__(self):
        return "<%s %r>" % (self.cls.name.split("/")[-1], self.f)


class JArr:
    __slots__ = ("t", "a")

    def __init__(self, t, n, fill=None):
        self.t = t
        self.a = [fill] * n


class JNative:
    """instance of a class that is not in the jars"""
    __slots__ = ("name", "v", "f")

    def __init__(self, name, v=None):
        self.name, self.v, self.f = name, v, {}


class ClassRef:
    def __init__(self, name):
        self.name = name


class PySet:
    """membership-only stand-in for fastutil LongSet / Long2LongOpenHashMap keys / the Illumina data holders (BarcodesMap additionally
    hands out its empty-drop list)"""

    def __init__(self, keys, empty_drops=None):
        self.s = set(int(k) for k in keys)
        self.empty_drops = empty_drops


class JdkHashSet:
    """java.util.HashSet = HashMap keys, modelled as far as the ITERATION ORDER goes (JDK 8+): table of 16 doubling past a load of 0.75,
    bucket = spread(hashCode) & (capacity - 1) with spread(h) = h ^ (h >>> 16), chains in insertion order, resize() splits a chain
    preserving relative order, treeifyBin() resizes instead of treeifying below 64 buckets (a treeified bin is flagged: its order is
    not modelled).  hashCode() / equals() are the element class's own bytecode."""

    def __init__(self, vm):
        self.vm, self.table, self.size, self.threshold, self.treeified = vm, None, 0, 0, False

    def _resize(self):
        if self.table is None:
            self.table, self.threshold = [[] for _ in range(16)], 12
            return
        old, ncap = self.table, len(self.table) * 2
        self.threshold *= 2
        self.table = [[] for _ in range(ncap)]
        for chain in old:
            for h, e in chain:
                self.table[h & (ncap - 1)].append((h, e))

    def add(self, e):
        hc = self.vm.j_hash(e) & 0xFFFFFFFF
        h = hc ^ (hc >> 16)
        if self.table is None:
            self._resize()
        chain = self.table[h & (len(self.table) - 1)]
        for h2, e2 in chain:
            if h2 == h and (e2 is e or self.vm.j_equals(e, e2)):
                return 0
        chain.append((h, e))
        if len(chain) >= 9:
            if len(self.table) < 64:
                self._resize()
            else:
                self.treeified = True
        self.size += 1
        if self.size > self.threshold:
            self._resize()
        return 1

    def items(self):
        return [e for chain in (self.table or []) for _, e in chain]


class JStream:
    """sequential java.util.stream pipeline: lazy per element (filter / map run element by element at the terminal operation), sorted()
    is a barrier with a stable sort"""

    def __init__(self, src, ops=()):
        self.src, self.ops = src, list(ops)

    def iterate(self, vm, limit=None):
        """element by element, like the JDK's sequential pipeline: the consumer of element k runs before the filter of element k + 1"""
        n = 0
        for x in self.src:
            keep = True
            for kind, f in self.ops:
                if kind == "filter":
                    if not vm.call_functional(f, [x]):
                        keep = False
                        break
                elif kind == "map":
                    x = vm.call_functional(f, [x])
            if keep:
                yield x
                n += 1
                if limit is not None and n >= limit:
                    break

    def run(self, vm, limit=None):
        return list(self.iterate(vm, limit))


class FastutilIntMap:
    """it.unimi.dsi.fastutil.ints.Int2ObjectOpenHashMap as far as the path uses it: put / get and the ITERATION ORDER of keys and
    entries (oracle/pyref.fastutil_key_order: the published open-addressing layout of fastutil 8.2.2, jar absent — the order is a
    restatement, which is why the vectors record the order that was used)."""

    def __init__(self):
        self.d, self.inserted = {}, []

    def put(self, k, v):
        if k not in self.d:
            self.inserted.append(k)
        self.d[k] = v

    def order(self):
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import pyref
        return pyref.fastutil_key_order(self.inserted)


def default_of(desc):
    c = desc[0]
    if c == "J":
        return L(0)
    if c in "ISBCZ":
        return 0
    if c in "FD":
        return D(0.0) if c == "D" else 0.0
    return None


def parse_desc(desc):
    """(list of argument type chars, return type char)"""
    args, i = [], 1
    while desc[i] != ")":
        c = desc[i]
        if c == "L":
            i = desc.index(";", i)
            args.append("L")
        elif c == "[":
            while desc[i] == "[":
                i += 1
            if desc[i] == "L":
                i = desc.index(";", i)
            args.append("[")
        else:
            args.append(c)
        i += 1
    return args, desc[i + 1]


class JClass:
    def __init__(self, vm, name, data):
        self.vm, self.name = vm, name
        cf = self.cf = jdis.ClassFile(data)
        self.super_name = cf.super
        self.ifaces = cf.ifaces
        self.statics, self.inst_fields, self.methods = {}, [], {}
        for acc, fname, desc, attrs in cf.fields:
            if acc & 8:
                v = default_of(desc)
                for an, d in attrs:
                    if an == "ConstantValue":
                        e = cf.cp[struct.unpack(">H", d)[0]]
                        v = L(e[1]) if e[0] == "Long" else (cf.utf(e[1]) if e[0] == "String" else e[1])
                self.statics[fname] = v
            else:
                self.inst_fields.append((fname, desc))
        for acc, mname, desc, attrs in cf.methods:
            code = None
            for an, d in attrs:
                if an == "Code":
                    r = jdis.R(d)
                    r.u2()
                    nloc = r.u2()
                    cl = r.u4()
                    code = (nloc, bytes(r.raw(cl)))
            self.methods[mname + desc] = (acc, code, parse_desc(desc))
        self.initialized = False
        self._cpc = {}
        self.bootstrap = []                      # BootstrapMethods: (method-handle cp index, [argument cp indices])
        for an, d in cf.attrs:
            if an == "BootstrapMethods":
                r = jdis.R(d)
                for _ in range(r.u2()):
                    ref = r.u2()
                    self.bootstrap.append((ref, [r.u2() for _ in range(r.u2())]))

    def is_subclass_of(self, name):
        c = self
        while c is not None:
            if c.name == name or name in c.ifaces:
                return True
            c = self.vm.load(c.super_name) if c.super_name else None
        return False


class VM:
    def __init__(self, jars):
        self.z = {}
        for j in jars:
            z = zipfile.ZipFile(j)
            for n in z.namelist():
                if n.endswith(".class"):
                    self.z.setdefault(n[:-6], (z, n))
        self.classes = {}
        self.n_insn = 0

    # ------------------------------------------------------------------------------------------------ class loading
    def load(self, name):
        c = self.classes.get(name)
        if c is None and name in self.z:
            z, n = self.z[name]
            c = self.classes[name] = JClass(self, name, z.read(n))
        return c

    def init_class(self, c):
        if c.initialized:
            return
        c.initialized = True
        if c.super_name:
            s = self.load(c.super_name)
            if s:
                self.init_class(s)
        if "<clinit>()V" in c.methods:
            self.run(c, "<clinit>()V", [])

    def new_object(self, c, init=True):
        """init=False: a bare field holder (no static initialiser, no constructor) for parameter objects whose classes pull in the whole
        application (XML binding, Runtime, logging) but of which the path only reads a few fields"""
        if init:
            self.init_class(c)
        o = JObj(c)
        k = c
        while k is not None:
            for fname, desc in k.inst_fields:
                o.f[fname] = default_of(desc)
            k = self.load(k.super_name) if k.super_name else None
        return o

    def find_method(self, c, key):
        while c is not None:
            m = c.methods.get(key)
            if m is not None and m[1] is not None:
                return c, m
            if not c.super_name:
                return None, None
            nxt = self.load(c.super_name)
            if nxt is None:
                return c.super_name, None              # continues in a class outside the jars
            c = nxt
        return None, None

    # ------------------------------------------------------------------------------------------------ public entry points
    def call_static(self, cls, name, desc, *args):
        c = self.load(cls)
        self.init_class(c)
        return self.run(c, name + desc, list(args))

    def call_virtual(self, obj, name, desc, *args):
        return self.invoke_virtual(obj.cls.name, name, desc, [obj] + list(args))

    def construct(self, cls, desc, *args):
        c = self.load(cls)
        o = self.new_object(c)
        self.run(c, "<init>" + desc, [o] + list(args))
        return o

    # ------------------------------------------------------------------------------------------------ invocation
    def invoke_virtual(self, decl_cls, name, desc, args):
        recv = args[0]
        if recv is None:
            raise JavaThrow("java/lang/NullPointerException", "%s.%s" % (decl_cls, name))
        if isinstance(recv, JObj):
            c, m = self.find_method(recv.cls, name + desc)
            if m is not None:
                return self.run(c, name + desc, args)
            return self.native(c if isinstance(c, str) else decl_cls, name, desc, args)
        return self.native(decl_cls, name, desc, args)

    def invoke_exact(self, cls, name, desc, args):
        c = self.load(cls)
        if c is None:
            return self.native(cls, name, desc, args)
        self.init_class(c)
        k, m = self.find_method(c, name + desc)
        if m is None:
            return self.native(k if isinstance(k, str) else cls, name, desc, args)
        return self.run(k, name + desc, args)

    def call_lambda(self, lam, args):
        cls, name, desc, kind = lam.v
        return self.invoke_exact(cls, name, desc, list(lam.f["captured"]) + list(args))

    def call_functional(self, f, args):
        """Function.apply / Predicate.test / Consumer.accept / ToIntFunction.applyAsInt on a lambda, a comparator key or a Python shim"""
        if callable(f):
            return f(*args)
        if f.name == "lambda":
            return self.call_lambda(f, args)
        if f.name == "pyfunc":
            return f.v(*args)
        raise NotImplementedError("functional object %s" % f.name)

    def new_container(self, supplier):
        """Supplier of a collector: a constructor reference (HashSet::new, ArrayList::new, Int2ObjectOpenHashMap::new, ConcurrentHashMap::new)"""
        target = supplier.v[0]
        if target == "java/util/HashSet":
            return JNative(target, JdkHashSet(self))
        if target in ("java/util/ArrayList", "java/util/LinkedList"):
            return JNative(target, [])
        if target.endswith("Int2ObjectOpenHashMap"):
            return JNative(target, FastutilIntMap())
        if target in ("java/util/concurrent/ConcurrentHashMap", "java/util/HashMap"):
            return JNative(target, {})
        raise NotImplementedError("collector supplier %s" % target)

    def collect(self, items, col):
        """Stream.collect(Collector) for the collectors the path uses"""
        kind = col.name if col is not None else "collector:toList"
        if kind == "collector:toList":
            return JNative("java/util/ArrayList", list(items))
        if kind == "collector:toSet":
            hs = JdkHashSet(self)
            for x in items:
                hs.add(x)
            return JNative("java/util/HashSet", hs)
        if kind == "collector:toCollection":
            c = self.new_container(col.v[0])
            for x in items:
                c.v.add(x) if isinstance(c.v, JdkHashSet) else c.v.append(x)
            return c
        if kind == "collector:mapping":
            return self.collect([self.call_functional(col.v[0], [x]) for x in items], col.v[1])
        if kind == "collector:toMap":                       # (keyMapper, valueMapper, mergeFunction, mapSupplier): Map.merge per element
            kf, vf, merge, sup = col.v
            c = self.new_container(sup)
            for x in items:
                k_, v_ = self.call_functional(kf, [x]), self.call_functional(vf, [x])
                if isinstance(c.v, FastutilIntMap):
                    c.v.put(k_, self.call_functional(merge, [c.v.d[k_], v_]) if k_ in c.v.d else v_)
                else:
                    c.v[k_] = self.call_functional(merge, [c.v[k_], v_]) if k_ in c.v else v_
            return c
        if kind == "collector:groupingBy":                  # (classifier, mapFactory, downstream)
            classifier, sup, down = col.v
            c = self.new_container(sup)
            groups = {}
            for x in items:
                groups.setdefault(self.call_functional(classifier, [x]), []).append(x)
            for k_, xs in groups.items():
                c.v[k_] = self.collect(xs, down)
            return c
        raise NotImplementedError(kind)

    def j_string(self, v):
        """String.valueOf(Object) as string concatenation applies it"""
        if v is None:
            return "null"
        if isinstance(v, str):
            return v
        if isinstance(v, JObj):
            c, m = self.find_method(v.cls, "toString()Ljava/lang/String;")
            if m is not None:
                return self.run(c, "toString()Ljava/lang/String;", [v])
            return v.cls.name.replace("/", ".") + "@0"
        if isinstance(v, float):
            return repr(float(v))
        if isinstance(v, JNative):
            if v.name == "java/lang/String":
                return v.v
            if v.name.endswith("Optional"):
                return "Optional[%s]" % self.j_string(v.v[0]) if v.v else "Optional.empty"
            return v.name.replace("/", ".") + "@0"
        return str(int(v))

    def j_hash(self, a):
        if isinstance(a, JObj):
            c, m = self.find_method(a.cls, "hashCode()I")
            if m is not None:
                return self.run(c, "hashCode()I", [a])
            return id(a) & 0x7FFFFFFF
        if isinstance(a, L):
            v = int(a) & 0xFFFFFFFFFFFFFFFF
            return i32(v ^ (v >> 32))                      # Long.hashCode
        if isinstance(a, int):
            return a
        return hash(a) & 0x7FFFFFFF

    def j_compare(self, a, b):
        c, m = self.find_method(a.cls, "compareTo(Ljava/lang/Object;)I")
        return self.run(c, "compareTo(Ljava/lang/Object;)I", [a, b])

    # ------------------------------------------------------------------------------------------------ natives (shims)
    def j_equals(self, a, b):
        if isinstance(a, JObj):
            c, m = self.find_method(a.cls, "equals(Ljava/lang/Object;)Z")
            if m is not None:
                return bool(self.run(c, "equals(Ljava/lang/Object;)Z", [a, b]))
            return a is b
        return a == b

    def native(self, cls, name, desc, args):
        key = name + desc
        a = args
        if name == "<init>":
            o = a[0]
            if cls == "java/lang/Object":
                return None
            if cls == "java/util/HashSet":
                hs = JdkHashSet(self)
                if isinstance(o, JNative):
                    o.v = hs
                else:
                    o.native = hs
                return None
            if cls in ("java/util/ArrayDeque", "java/util/ArrayList", "java/util/LinkedList"):
                (o if isinstance(o, JNative) else o).__setattr__("v" if isinstance(o, JNative) else "native", [])
                return None
            if cls in ("java/util/concurrent/ConcurrentHashMap", "java/util/HashMap", "java/util/TreeMap"):
                o.v = {}
                return None
            if cls == "java/util/concurrent/atomic/AtomicInteger":
                o.v = [a[1] if len(a) > 1 else 0]
                return None
            if cls.endswith("IntHashSet") or cls.endswith("LongHashSet"):
                o.v = set()
                return None
            if cls == "java/lang/String":
                o.v = "".join(chr(c) for c in a[1].a)
                return None
            if cls == "java/lang/StringBuilder":
                o.v = [a[1]] if len(a) > 1 and isinstance(a[1], str) else []
                return None
            if cls in ("java/lang/AssertionError", "java/lang/IllegalArgumentException", "java/lang/Enum"):
                if cls == "java/lang/Enum":
                    o.f["$name"], o.f["$ordinal"] = a[1], a[2]
                else:
                    o.v = a[1] if len(a) > 1 else None
                return None
            if cls.endswith("tuple/MutableTriple"):
                o.f["left"], o.f["middle"], o.f["right"] = a[1], a[2], a[3]
                return None
            # any other class outside the jars: an opaque object (formatters, loggers, ... created by static initialisers); calling a
            # method on it still fails loudly
            if isinstance(o, JNative):
                o.v = tuple(a[1:])
            return None
        store = None
        if a:
            store = a[0].v if isinstance(a[0], JNative) else (a[0].native if isinstance(a[0], JObj) else None)
        if cls == "htsjdk/samtools/fastq/FastqRecord" and name in getattr(self, "fastq_fields", {}):
            return self.fastq_fields[name]                 # the superclass of FastqRecordExt is outside the jars: the driver supplies its getters
        if name == "clone" and a and isinstance(a[0], JArr):
            c_ = JArr(a[0].t, 0)
            c_.a = list(a[0].a)
            return c_
        if a and isinstance(a[0], JNative) and a[0].name == "logger":
            return None                                    # log4j / java.util.logging calls are dropped
        if cls in ("java/lang/Long", "java/lang/Integer", "java/lang/Boolean", "java/lang/Byte", "java/lang/Short"):
            if name == "valueOf":
                return a[0]
            if name == "byteValue":
                v = a[0] & 0xFF
                return v - 256 if v & 0x80 else v
            if name.endswith("Value") and name != "shortValue":
                return a[0]
        if cls == "java/lang/Math":
            return {"abs": lambda: abs(a[0]), "max": lambda: max(a[0], a[1]), "min": lambda: min(a[0], a[1])}[name]()
        if cls == "java/lang/Class":
            if name == "desiredAssertionStatus":
                return 0                                   # java runs without -ea
            if name == "getName":
                return a[0].name.replace("/", ".")
        if cls in ("org/apache/logging/log4j/LogManager", "java/util/logging/Logger") and name == "getLogger":
            return JNative("logger")
        if cls in ("java/util/Deque", "java/util/ArrayDeque"):
            if name == "add":
                store.append(a[1])
                return 1
            if name == "pollLast":
                return store.pop() if store else None
            if name == "isEmpty":
                return int(not store)
        if cls in ("java/util/List", "java/util/ArrayList", "java/util/Collection", "java/util/LinkedList"):
            if name == "add" and len(a) == 3:
                store.insert(a[1], a[2])
                return None
            if name == "toArray":
                arr = JArr("L", 0, None)
                arr.a = list(store)
                return arr
            if name == "add":
                store.append(a[1])
                return 1
            if name == "isEmpty":
                return int(not store)
            if name == "size":
                return len(store)
            if name == "get":
                if not 0 <= a[1] < len(store):
                    raise JavaThrow("java/lang/IndexOutOfBoundsException")
                return store[a[1]]
            if name == "addAll":
                src = a[1].v if isinstance(a[1], JNative) else a[1].native
                store.extend(src)
                return int(bool(src))
            if name == "forEach":
                for x in list(store):
                    self.call_functional(a[1], [x])
                return None
            if name == "stream":
                return JNative("java/util/stream/Stream", JStream(list(store)))
        if isinstance(store, JdkHashSet):
            if name == "add":
                return store.add(a[1])
            if name == "addAll":
                src = a[1].native if isinstance(a[1], JObj) else a[1].v
                ch = 0
                for e in (src.items() if isinstance(src, JdkHashSet) else list(src)):
                    ch |= store.add(e)
                return ch
            if name == "isEmpty":
                return int(store.size == 0)
            if name == "size":
                return store.size
            if name == "contains":
                return int(any(self.j_equals(a[1], e) for e in store.items()))
            if name == "stream":
                return JNative("java/util/stream/Stream", JStream(store.items()))
        if a and isinstance(a[0], JNative) and a[0].name in ("lambda", "pyfunc") and name in ("apply", "test", "accept", "applyAsInt", "get", "compare"):
            return self.call_functional(a[0], a[1:])
        if cls == "java/util/stream/IntStream" and name == "rangeClosed":
            return JNative("java/util/stream/Stream", JStream(list(range(a[0], a[1] + 1))))
        if cls in ("java/util/stream/Stream", "java/util/stream/IntStream") and isinstance(a[0].v, JStream):
            st_ = a[0].v
            if name == "boxed":
                return a[0]
            if name in ("filter", "map") or (name == "mapToObj"):
                return JNative("java/util/stream/Stream", JStream(st_.src, st_.ops + [("map" if name == "mapToObj" else name, a[1])]))
            if name == "sorted":
                import functools
                items = st_.run(self)
                if len(a) > 1 and isinstance(a[1], JNative) and a[1].name == "comparator":        # Comparator.comparingInt(key)
                    items.sort(key=lambda x: self.call_functional(a[1].v, [x]))
                elif len(a) > 1:                            # a Comparator: lambda or an object whose class is in the jars
                    cmpf = (lambda x, y: self.call_functional(a[1], [x, y])) if isinstance(a[1], JNative) else \
                           (lambda x, y: self.invoke_virtual(a[1].cls.name, "compare", "(Ljava/lang/Object;Ljava/lang/Object;)I", [a[1], x, y]))
                    items.sort(key=functools.cmp_to_key(cmpf))
                else:                                       # natural order: the element's own compareTo (stable, like the JDK's TimSort)
                    items.sort(key=functools.cmp_to_key(lambda x, y: self.j_compare(x, y)))
                return JNative("java/util/stream/Stream", JStream(items))
            if name == "skip":
                if int(a[1]) < 0:                          # ReferencePipeline.skip: IllegalArgumentException(Long.toString(n))
                    raise JavaThrow("java/lang/IllegalArgumentException", str(int(a[1])))
                return JNative("java/util/stream/Stream", JStream(st_.run(self)[int(a[1]):]))
            if name == "mapToInt":
                return JNative("java/util/stream/Stream", JStream(st_.src, st_.ops + [("map", a[1])]))
            if name == "limit":
                return JNative("java/util/stream/Stream", JStream(st_.run(self, limit=int(a[1]))))
            if name == "count":
                return L(len(st_.run(self)))
            if name == "sum":
                return i32(sum(st_.run(self)))
            if name in ("min", "max") and len(a) == 1:      # IntStream.min / max
                v_ = st_.run(self)
                return JNative("java/util/OptionalInt", ((min(v_) if name == "min" else max(v_)),) if v_ else ())
            if name == "average":
                v_ = st_.run(self)
                return JNative("java/util/OptionalDouble", (D(sum(v_) / len(v_)),) if v_ else ())
            if name == "distinct":                         # LinkedHashSet semantics: first occurrence wins, equals() decides
                out_ = []
                for x in st_.run(self):
                    if not any(self.j_equals(x, y) for y in out_):
                        out_.append(x)
                return JNative("java/util/stream/Stream", JStream(out_))
            if name in ("parallel", "sequential"):          # the interpreter runs every pipeline sequentially
                return a[0]
            if name == "collect":
                return self.collect(st_.run(self), a[1] if len(a) > 1 else None)
            if name == "forEach":                           # lazily: a consumer that changes what the next filter reads is seen by it
                for x in st_.iterate(self):
                    self.call_functional(a[1], [x])
                return None
            if name == "findFirst":
                r = st_.run(self, limit=1)
                return JNative("java/util/Optional", (r[0],) if r else ())
            if name == "max":                               # Stream.max(comparator) = reduce(BinaryOperator.maxBy): compare(a, b) >= 0 ? a : b,
                items = st_.run(self)                       # the FIRST of equal maxima stays
                if not items:
                    return JNative("java/util/Optional", ())
                best = items[0]
                for x in items[1:]:
                    if not self.call_functional(a[1].v, [best]) >= self.call_functional(a[1].v, [x]):
                        best = x
                return JNative("java/util/Optional", (best,))
        if cls.endswith("lang3/ArrayUtils") and name == "toPrimitive":
            arr = JArr("B", 0, 0)
            arr.a = list(a[0].a)
            return arr
        if cls == "java/util/OptionalInt":
            if name == "orElse":
                return a[0].v[0] if a[0].v else a[1]
            if name == "getAsInt":
                return a[0].v[0]
        if cls == "java/util/OptionalDouble" and name == "getAsDouble":
            if not a[0].v:
                raise JavaThrow("java/util/NoSuchElementException")
            return a[0].v[0]
        if cls == "java/text/DecimalFormat" and name == "format":
            # DecimalFormat(pattern).format(Number): only the two patterns of the path ("##.#": at most one fraction digit; "###,###,###,###":
            # grouping), RoundingMode.HALF_EVEN on the exact binary value (JDK >= 8)
            from decimal import Decimal, ROUND_HALF_EVEN
            pat, v_ = a[0].v[0], a[1]
            if "," in pat:
                return format(int(v_), ",")
            d_ = Decimal(float(v_)).quantize(Decimal("0.1"), rounding=ROUND_HALF_EVEN)
            t_ = format(d_, "f")
            return t_[:-2] if t_.endswith(".0") else t_
        if cls == "java/lang/String" and name == "chars" and isinstance(a[0], str):
            return JNative("java/util/stream/Stream", JStream([ord(ch) for ch in a[0]]))
        if cls == "java/util/stream/Collectors" and name == "toList":
            return JNative("collector:toList")
        if cls == "java/util/stream/Collectors" and name in ("toCollection", "toSet", "toMap", "mapping", "groupingBy", "groupingByConcurrent"):
            return JNative("collector:" + name.replace("Concurrent", ""), tuple(a))
        if cls == "java/util/Comparator" and name == "comparing":
            return JNative("comparator", a[0])
        if a and isinstance(a[0], JNative) and isinstance(a[0].v, FastutilIntMap):
            fm = a[0].v
            if name == "keySet":
                return JNative("pylist", list(fm.order()))
            if name == "int2ObjectEntrySet":
                return JNative("pylist", [JNative("entry", (k_, fm.d[k_])) for k_ in fm.order()])
            if name == "get":
                return fm.d.get(a[1])
            if name == "size":
                return len(fm.d)
        if a and isinstance(a[0], JNative) and a[0].name == "pylist" and name in ("stream", "parallelStream"):
            return JNative("java/util/stream/Stream", JStream(list(a[0].v)))
        if a and isinstance(a[0], JNative) and a[0].name == "entry" and name == "getIntKey":
            return a[0].v[0]
        if isinstance(store, dict) and name in ("isEmpty", "values", "size") and len(a) == 1:
            if name == "values":
                return JNative("java/util/ArrayList", list(store.values()))
            return int(not store) if name == "isEmpty" else len(store)
        if cls == "java/util/Comparator" and name == "comparingInt":
            return JNative("comparator", a[0])
        if name == "entrySet" and isinstance(store, dict):      # TreeMap: ascending keys (the only ordered map on the path)
            return JNative("java/util/ArrayList", [JNative("entry", (k_, store[k_])) for k_ in sorted(store)])
        if a and isinstance(a[0], JNative) and a[0].name == "entry" and name in ("getKey", "getValue"):
            return a[0].v[0 if name == "getKey" else 1]
        if cls == "java/util/Map$Entry" and name == "comparingByKey":
            return JNative("comparator", lambda e_: e_.v[0])
        if cls in ("java/util/Map", "java/util/concurrent/ConcurrentHashMap", "java/util/HashMap", "java/util/TreeMap") and isinstance(store, dict):
            k = a[1]
            if name == "get":
                return store.get(k)
            if name == "put":
                old = store.get(k)
                store[k] = a[2]
                return old
            if name == "putIfAbsent":
                if k in store and store[k] is not None:
                    return store[k]
                store[k] = a[2]
                return None
            if name == "containsKey":
                return int(k in store)
        if cls == "java/util/concurrent/atomic/AtomicInteger":
            if name == "addAndGet":
                store[0] += a[1]
                return store[0]
            if name in ("incrementAndGet", "getAndIncrement"):
                store[0] += 1
                return store[0] if name == "incrementAndGet" else store[0] - 1
            if name == "get":
                return store[0]
        if cls.endswith("fastutil/longs/Long2ObjectMap") and isinstance(store, dict):
            if name == "containsKey":
                return int(int(a[1]) in store)
            if name == "get":
                return store.get(int(a[1]))
            if name == "put":
                old_ = store.get(int(a[1]))
                store[int(a[1])] = a[2]
                return old_
        if cls.endswith("Long2ObjectOpenHashMap") or cls.endswith("BarcodesMapForBCfinding"):
            if name == "keySet":
                return PySet(a[0].native.keys())
            if name == "get":
                return a[0].native.get(int(a[1]))
        if cls in ("java/util/HashSet", "java/util/Set", "java/util/AbstractCollection", "java/util/AbstractSet"):
            if isinstance(a[0], PySet):
                if name == "contains":
                    return int(int(a[1]) in a[0].s)
                if name == "isEmpty":
                    return int(not a[0].s)
            if name == "add":                              # HashSet.add: equals() decides (hashCode consistency is the class's business)
                for e in store:
                    if self.j_equals(e, a[1]):
                        return 0
                store.append(a[1])
                return 1
            if name == "isEmpty":
                return int(not store)
            if name == "size":
                return len(store)
            if name == "contains":
                return int(any(self.j_equals(e, a[1]) for e in store))
        if cls in ("java/util/Optional", "com/google/common/base/Optional") and name == "orElse":
            return a[0].v[0] if a[0].v else a[1]
        if cls in ("java/util/Optional", "com/google/common/base/Optional") and name in ("get", "isPresent"):
            if name == "isPresent":
                return int(len(a[0].v) == 1)
            if not a[0].v:
                raise JavaThrow("java/util/NoSuchElementException")
            return a[0].v[0]
        if cls in ("java/lang/String", "java/lang/CharSequence", "java/lang/Object") and isinstance(a[0] if a else None, str) and name in ("indexOf", "lastIndexOf", "split", "equals", "isEmpty", "trim",
                                                                                          "startsWith", "endsWith", "contains", "toString", "hashCode"):
            t = a[0]
            arg = a[1] if len(a) > 1 else None
            if name in ("indexOf", "lastIndexOf"):
                needle = chr(arg) if isinstance(arg, int) else arg
                return (t.find(needle, a[2]) if len(a) > 2 else t.find(needle)) if name == "indexOf" else t.rfind(needle)
            if name == "split":
                arr = JArr("L", 0, None)
                arr.a = t.split(arg)
                while arr.a and arr.a[-1] == "":
                    arr.a.pop()
                return arr
            if name == "equals":
                return int(t == arg)
            if name == "isEmpty":
                return int(not t)
            if name == "trim":
                return t.strip()
            if name == "startsWith":
                return int(t.startswith(arg))
            if name == "endsWith":
                return int(t.endswith(arg))
            if name == "contains":
                return int(arg in t)
            if name == "toString":
                return t
            if name == "hashCode":
                h = 0
                for ch in t:
                    h = i32(31 * h + ord(ch))
                return h
        if cls in ("java/lang/Integer", "java/lang/Float", "java/lang/Long") and name in ("parseInt", "parseFloat", "parseLong"):
            try:
                if name == "parseFloat":
                    return float(np_float32(a[0]))
                radix = a[1] if len(a) > 1 else 10
                t = a[0]
                digits = t[1:] if t[:1] in "+-" else t
                if not digits or any(not ch.isalnum() or int(ch, 36) >= radix for ch in digits):      # Character.digit(ch, radix) < 0
                    raise ValueError
                v = int(t, radix)
                if not -(1 << 31) <= v < (1 << 31) and name == "parseInt":
                    raise ValueError
                return v
            except ValueError:
                raise JavaThrow("java/lang/NumberFormatException", a[0])
        if cls == "java/lang/Float" and name == "valueOf":
            return a[0]
        if cls == "java/lang/Integer" and name == "toString" and len(a) == 2:
            n_, r_ = a
            out = ""
            m_ = abs(n_)
            while True:
                out = "0123456789abcdefghijklmnopqrstuvwxyz"[m_ % r_] + out
                m_ //= r_
                if m_ == 0:
                    break
            return ("-" if n_ < 0 else "") + out
        if cls == "java/lang/String" and name == "valueOf":
            return str(int(a[0])) if isinstance(a[0], int) else str(a[0])
        if cls == "java/lang/Integer" and name == "shortValue":
            v = a[0] & 0xFFFF
            return v - 0x10000 if v & 0x8000 else v
        if cls in ("java/lang/String", "java/lang/CharSequence") and name in ("substring", "subSequence") and isinstance(a[0], str):
            b, e = a[1], (a[2] if len(a) > 2 else len(a[0]))
            if b < 0 or e > len(a[0]) or b > e:
                raise JavaThrow("java/lang/StringIndexOutOfBoundsException", "begin %d, end %d, length %d" % (b, e, len(a[0])))
            return a[0][b:e]
        if cls == "com/google/common/base/Optional" and name in ("absent", "of", "fromNullable"):
            return JNative(cls, () if name == "absent" or a[0] is None else (a[0],))
        if cls == "java/util/Optional":
            if name == "empty":
                return JNative("java/util/Optional", ())
            if name == "of":
                return JNative("java/util/Optional", (a[0],))
            if name == "isPresent":
                return int(len(a[0].v) == 1)
            if name == "isEmpty":
                return int(len(a[0].v) == 0)
            if name == "get":
                if not a[0].v:
                    raise JavaThrow("java/util/NoSuchElementException")
                return a[0].v[0]
        if cls.endswith("IntHashSet") or cls.endswith("LongHashSet"):
            if name == "add":
                n0 = len(store)
                store.add(int(a[1]))
                return int(len(store) != n0)
            if name == "contains":
                return int(int(a[1]) in store)
        if isinstance(a[0] if a else None, PySet) and name in ("contains", "containsKey"):
            return int(int(a[1]) in a[0].s)
        if isinstance(a[0] if a else None, PySet) and name == "size":
            return len(a[0].s)
        if isinstance(a[0] if a else None, PySet) and name == "isEmpty":
            return int(not a[0].s)
        if isinstance(a[0] if a else None, PySet) and name == "getEmptyDropBarcodes":
            return a[0].empty_drops
        if a and isinstance(a[0], JNative) and a[0].name == "java/util/AbstractMap$SimpleEntry":
            if name == "getKey":
                return a[0].v[0]
            if name == "getValue":
                return a[0].v[1]
            if name == "setValue":
                old_ = a[0].v[1]
                a[0].v = (a[0].v[0], a[1])
                return old_
        if cls == "java/util/concurrent/atomic/AtomicBoolean":
            if name == "get":
                return int(bool(a[0].v[0])) if a[0].v else 0
            if name == "set":
                a[0].v = (a[1],)
                return None
        if cls in ("java/lang/String", "java/lang/CharSequence"):
            s = a[0].v if isinstance(a[0], JNative) else a[0]
            if isinstance(s, str):
                if name == "length":
                    return len(s)
                if name == "charAt":
                    if not 0 <= a[1] < len(s):
                        raise JavaThrow("java/lang/StringIndexOutOfBoundsException")
                    return ord(s[a[1]])
                if name == "toCharArray":
                    arr = JArr("C", len(s), 0)
                    arr.a = [ord(ch) for ch in s]
                    return arr
        if cls == "java/lang/StringBuilder":
            if name == "append":
                store.append(chr(a[1]) if desc.startswith("(C)") else self.j_string(a[1]))
                return a[0]
            if name == "toString":
                return "".join(store)
        if cls == "java/lang/System" and name == "arraycopy":
            src, sp, dst, dp, n = a
            if sp < 0 or dp < 0 or n < 0 or sp + n > len(src.a) or dp + n > len(dst.a):
                raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException", "arraycopy")
            dst.a[dp:dp + n] = src.a[sp:sp + n]
            return None
        if cls == "java/util/Arrays":
            if name == "copyOf":
                r = JArr(a[0].t, a[1], 0)
                r.a[:min(a[1], len(a[0].a))] = a[0].a[:a[1]]
                return r
            if name == "fill":
                if len(a) == 2:
                    a[0].a[:] = [a[1]] * len(a[0].a)
                else:
                    a[0].a[a[1]:a[2]] = [a[3]] * (a[2] - a[1])
                return None
            if name == "equals":
                return int(a[0] is a[1] or (a[0] is not None and a[1] is not None and a[0].a == a[1].a))
            if name == "hashCode":
                h = 1
                for b in a[0].a:
                    h = i32(31 * h + b)
                return h
        if cls == "java/lang/Object":
            if name == "hashCode":
                return id(a[0]) & 0x7FFFFFFF
            if name == "equals":
                return int(a[0] is a[1])
            if name == "getClass":
                return ClassRef(a[0].cls.name if isinstance(a[0], JObj) else a[0].name)
        if cls == "java/lang/Enum":
            if name == "ordinal":
                return a[0].f["$ordinal"]
            if name == "name":
                return a[0].f["$name"]
        if cls.endswith("tuple/MutableTriple") or cls.endswith("tuple/Triple"):
            part = {"Left": "left", "Middle": "middle", "Right": "right"}[name[3:]]
            if name.startswith("get"):
                return a[0].f[part]
            a[0].f[part] = a[1]
            return None
        if name == "stream" and isinstance(a[0], JNative) and isinstance(a[0].v, list):
            return JNative("java/util/stream/Stream", JStream(list(a[0].v)))
        if cls == "java/util/EnumSet":
            if name in ("of", "allOf", "noneOf"):
                return JNative("java/util/EnumSet", sorted([x for x in a if isinstance(x, JObj)], key=lambda e: e.f["$ordinal"]))
            if name == "iterator":
                return JNative("java/util/Iterator", [list(a[0].v), 0])
            if name == "size":
                return len(a[0].v)
            if name == "contains":
                return int(any(e is a[1] for e in a[0].v))
        if cls == "java/util/Iterator":
            if name == "hasNext":
                return int(a[0].v[1] < len(a[0].v[0]))
            if name == "next":
                a[0].v[1] += 1
                return a[0].v[0][a[0].v[1] - 1]
        if cls == "java/util/Objects" and name == "requireNonNull":
            if a[0] is None:
                raise JavaThrow("java/lang/NullPointerException")
            return a[0]
        raise NotImplementedError("native %s.%s%s" % (cls, name, desc))

    # ------------------------------------------------------------------------------------------------ the interpreter
    def run(self, c, key, args):
        acc, code, (atypes, rtype) = c.methods[key]
        if code is None:
            raise NotImplementedError("abstract/native %s.%s" % (c.name, key))
        nloc, bc = code
        loc = [None] * (nloc + 2)
        i = 0
        for v in args:
            loc[i] = v
            i += 2 if isinstance(v, (L, D)) else 1
        cf, cp = c.cf, c.cf.cp
        st = []
        push, pop = st.append, st.pop
        pc = 0
        u2 = lambda p: (bc[p] << 8) | bc[p + 1]
        s2 = lambda p: ((bc[p] << 8) | bc[p + 1]) - (0x10000 if bc[p] & 0x80 else 0)
        ninsn = 0
        while True:
            op = bc[pc]
            ninsn += 1
            if op == 0x2a:                                  # aload_0
                push(loc[0]); pc += 1
            elif 0x1a <= op <= 0x2d:                        # iload_n lload_n fload_n dload_n aload_n
                push(loc[(op - 0x1a) & 3]); pc += 1
            elif op in (0x15, 0x16, 0x17, 0x18, 0x19):      # iload lload fload dload aload
                push(loc[bc[pc + 1]]); pc += 2
            elif 0x3b <= op <= 0x4e:                        # istore_n .. astore_n
                loc[(op - 0x3b) & 3] = pop(); pc += 1
            elif op in (0x36, 0x37, 0x38, 0x39, 0x3a):
                loc[bc[pc + 1]] = pop(); pc += 2
            elif op == 0xb4:                                # getfield
                e = cp[u2(pc + 1)]
                o = pop()
                if o is None:
                    raise JavaThrow("java/lang/NullPointerException", "getfield")
                push(o.f[cf.utf(cp[e[2]][1])]); pc += 3
            elif op == 0xb5:                                # putfield
                e = cp[u2(pc + 1)]
                v = pop(); o = pop()
                if o is None:
                    raise JavaThrow("java/lang/NullPointerException", "putfield")
                o.f[cf.utf(cp[e[2]][1])] = v; pc += 3
            elif 0x02 <= op <= 0x08:                        # iconst_m1..5
                push(op - 3); pc += 1
            elif op == 0x10:
                push(bc[pc + 1] - (256 if bc[pc + 1] & 0x80 else 0)); pc += 2
            elif op == 0x11:
                push(s2(pc + 1)); pc += 3
            elif op == 0x01:
                push(None); pc += 1
            elif op in (0x09, 0x0a):
                push(L(op - 9)); pc += 1
            elif op in (0x12, 0x13, 0x14):                  # ldc ldc_w ldc2_w
                idx = bc[pc + 1] if op == 0x12 else u2(pc + 1)
                e = cp[idx]
                t = e[0]
                push(L(e[1]) if t == "Long" else D(e[1]) if t == "Double" else cf.utf(e[1]) if t == "String" else
                     ClassRef(cf.utf(e[1])) if t == "Class" else e[1])
                pc += 2 if op == 0x12 else 3
            # ---- arithmetic ----
            elif op == 0x60:
                b = pop(); push(i32(pop() + b)); pc += 1
            elif op == 0x64:
                b = pop(); push(i32(pop() - b)); pc += 1
            elif op == 0x68:
                b = pop(); push(i32(pop() * b)); pc += 1
            elif op == 0x6c or op == 0x70:                  # idiv irem
                b = pop(); a = pop()
                if b == 0:
                    raise JavaThrow("java/lang/ArithmeticException")
                q = abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
                push(i32(q) if op == 0x6c else i32(a - q * b)); pc += 1
            elif op == 0x6d or op == 0x71:                  # ldiv lrem (truncating, like idiv)
                b = pop(); a = pop()
                if b == 0:
                    raise JavaThrow("java/lang/ArithmeticException")
                q = abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
                push(i64(q) if op == 0x6d else i64(a - q * b)); pc += 1
            elif op == 0x74:
                push(i32(-pop())); pc += 1
            elif op == 0x78:
                b = pop(); push(i32(pop() << (b & 31))); pc += 1
            elif op == 0x7a:
                b = pop(); push(pop() >> (b & 31)); pc += 1
            elif op == 0x7c:
                b = pop(); push(i32((pop() & 0xFFFFFFFF) >> (b & 31))); pc += 1
            elif op == 0x7e:
                b = pop(); push(pop() & b); pc += 1
            elif op == 0x80:
                b = pop(); push(pop() | b); pc += 1
            elif op == 0x82:
                b = pop(); push(pop() ^ b); pc += 1
            elif op == 0x61:
                b = pop(); push(i64(pop() + b)); pc += 1
            elif op == 0x65:
                b = pop(); push(i64(pop() - b)); pc += 1
            elif op == 0x69:
                b = pop(); push(i64(pop() * b)); pc += 1
            elif op == 0x75:
                push(i64(-pop())); pc += 1
            elif op == 0x79:                                # lshl: shift count is an int, masked with 63
                b = pop(); push(i64(pop() << (b & 63))); pc += 1
            elif op == 0x7b:
                b = pop(); push(L(pop() >> (b & 63))); pc += 1
            elif op == 0x7d:
                b = pop(); push(i64((pop() & 0xFFFFFFFFFFFFFFFF) >> (b & 63))); pc += 1
            elif op == 0x7f:
                b = pop(); push(L(pop() & b)); pc += 1
            elif op == 0x81:
                b = pop(); push(L(pop() | b)); pc += 1
            elif op == 0x83:
                b = pop(); push(L(pop() ^ b)); pc += 1
            elif op == 0x84:                                # iinc
                loc[bc[pc + 1]] = i32(loc[bc[pc + 1]] + (bc[pc + 2] - (256 if bc[pc + 2] & 0x80 else 0))); pc += 3
            elif op == 0x85:
                push(L(pop())); pc += 1
            elif op == 0x90:                                # d2f
                push(np_float32(pop())); pc += 1
            elif op == 0x8d:                                # f2d
                push(D(pop())); pc += 1
            elif op == 0x86 or op == 0x87:                  # i2f i2d
                v = pop(); push(np_float32(v) if op == 0x86 else D(v)); pc += 1
            elif op == 0x88:
                push(i32(pop())); pc += 1
            elif op == 0x91:
                v = pop() & 0xFF; push(v - 256 if v & 0x80 else v); pc += 1
            elif op == 0x92:
                push(pop() & 0xFFFF); pc += 1
            elif op == 0x93:
                v = pop() & 0xFFFF; push(v - 0x10000 if v & 0x8000 else v); pc += 1
            elif op == 0x94:
                b = pop(); a = pop(); push((a > b) - (a < b)); pc += 1
            elif op in (0x0e, 0x0f):                        # dconst_0 dconst_1
                push(D(op - 0x0e)); pc += 1
            elif op in (0x0b, 0x0c, 0x0d):                  # fconst_0 .. 2
                push(float(op - 0x0b)); pc += 1
            elif op in (0x95, 0x96, 0x97, 0x98):            # fcmpl fcmpg dcmpl dcmpg (NaN: -1 for the l forms, +1 for the g forms)
                b = pop(); a = pop()
                if a != a or b != b:
                    push(-1 if op in (0x95, 0x97) else 1)
                else:
                    push((a > b) - (a < b))
                pc += 1
            elif op in (0x63, 0x67, 0x6b, 0x6f):            # dadd dsub dmul ddiv
                b = float(pop()); a = float(pop())
                if op == 0x6f:
                    r_ = (a / b) if b != 0.0 else (float("nan") if a == 0.0 or a != a else float("inf") * (1 if (a > 0) == (str(b)[0] != "-") else -1))
                else:
                    r_ = a + b if op == 0x63 else a - b if op == 0x67 else a * b
                push(D(r_)); pc += 1
            elif op in (0x8e, 0x8b):                        # d2i f2i: truncation toward zero, saturating, NaN -> 0
                v = float(pop())
                push(0 if v != v else max(-(1 << 31), min((1 << 31) - 1, int(v)))); pc += 1
            elif op == 0x8f:                                # d2l
                v = float(pop())
                push(L(0 if v != v else max(-(1 << 63), min((1 << 63) - 1, int(v))))); pc += 1
            elif op == 0x8a:                                # l2d
                push(D(float(pop()))); pc += 1
            # ---- branches ----
            elif 0x99 <= op <= 0x9e:
                v = pop()
                t = (v == 0, v != 0, v < 0, v >= 0, v > 0, v <= 0)[op - 0x99]
                pc += s2(pc + 1) if t else 3
            elif 0x9f <= op <= 0xa4:
                b = pop(); a = pop()
                t = (a == b, a != b, a < b, a >= b, a > b, a <= b)[op - 0x9f]
                pc += s2(pc + 1) if t else 3
            elif op == 0xa5 or op == 0xa6:
                b = pop(); a = pop()
                pc += s2(pc + 1) if ((a is b) == (op == 0xa5)) else 3
            elif op == 0xa7:
                pc += s2(pc + 1)
            elif op == 0xc6 or op == 0xc7:
                v = pop()
                pc += s2(pc + 1) if ((v is None) == (op == 0xc6)) else 3
            # ---- arrays ----
            elif op in (0x2e, 0x2f, 0x32, 0x33, 0x34, 0x35, 0x30, 0x31):     # xaload
                i = pop(); arr = pop()
                if arr is None:
                    raise JavaThrow("java/lang/NullPointerException", "array load")
                if not 0 <= i < len(arr.a):
                    raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException", "index %d length %d" % (i, len(arr.a)))
                push(arr.a[i]); pc += 1
            elif op in (0x4f, 0x50, 0x53, 0x54, 0x55, 0x56, 0x51, 0x52):     # xastore
                v = pop(); i = pop(); arr = pop()
                if not 0 <= i < len(arr.a):
                    raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException", "index %d length %d" % (i, len(arr.a)))
                if op == 0x54:                                               # bastore truncates to byte
                    v &= 0xFF
                    v = v - 256 if v & 0x80 else v
                elif op == 0x55:
                    v &= 0xFFFF
                elif op == 0x56:
                    v &= 0xFFFF
                    v = v - 0x10000 if v & 0x8000 else v
                arr.a[i] = v; pc += 1
            elif op == 0xbc:                                # newarray
                n = pop()
                t = bc[pc + 1]
                push(JArr("J" if t == 11 else "D" if t == 7 else "I", n, L(0) if t == 11 else D(0.0) if t == 7 else 0)); pc += 2
            elif op == 0xbd:
                push(JArr("L", pop(), None)); pc += 3
            elif op == 0xc5:                                # multianewarray
                dims = bc[pc + 3]
                ns = [pop() for _ in range(dims)][::-1]
                desc = cf.utf(cp[u2(pc + 1)][1])
                leaf = desc[dims:]

                def mk(k):
                    if k == dims - 1:
                        return JArr(leaf[0], ns[k], default_of(leaf))
                    arr = JArr("L", ns[k], None)
                    arr.a = [mk(k + 1) for _ in range(ns[k])]
                    return arr
                push(mk(0)); pc += 4
            elif op == 0xbe:
                arr = pop()
                if arr is None:
                    raise JavaThrow("java/lang/NullPointerException", "arraylength")
                push(len(arr.a)); pc += 1
            # ---- stack ----
            elif op == 0x57:
                pop(); pc += 1
            elif op == 0x58:
                if not isinstance(pop(), (L, D)):
                    pop()
                pc += 1
            elif op == 0x59:
                push(st[-1]); pc += 1
            elif op == 0x5c:                                # dup2
                if isinstance(st[-1], (L, D)):
                    push(st[-1])
                else:
                    st.extend(st[-2:])
                pc += 1
            elif op == 0x5a:                                # dup_x1
                v = st[-1]; st.insert(-2, v); pc += 1
            # ---- statics, objects, calls ----
            elif op == 0xb2 or op == 0xb3:
                e = cp[u2(pc + 1)]
                cn = cf.utf(cp[e[1]][1])
                fn = cf.utf(cp[e[2]][1])
                k = self.load(cn)
                if k is None:
                    ext = {("java/lang/Boolean", "TRUE"): 1, ("java/lang/Boolean", "FALSE"): 0, ("java/lang/Integer", "MAX_VALUE"): 0x7FFFFFFF,
                           ("java/lang/Integer", "MIN_VALUE"): -0x80000000, ("java/lang/Long", "MAX_VALUE"): L(0x7FFFFFFFFFFFFFFF)}
                    if op == 0xb2 and (cn, fn) in ext:
                        push(ext[(cn, fn)]); pc += 3
                        continue
                    raise NotImplementedError("static field %s.%s" % (cn, fn))
                self.init_class(k)
                while fn not in k.statics:
                    k = self.load(k.super_name)
                if op == 0xb2:
                    push(k.statics[fn])
                else:
                    k.statics[fn] = pop()
                pc += 3
            elif op in (0xb6, 0xb7, 0xb8, 0xb9):
                e = cp[u2(pc + 1)]
                cn = cf.utf(cp[e[1]][1])
                mn, md = cf.nat(e[2])
                at, rt = c._cpc.get(md) or c._cpc.setdefault(md, parse_desc(md))
                n = len(at) + (op != 0xb8)
                cargs = st[len(st) - n:] if n else []
                del st[len(st) - n:]
                if op == 0xb8:
                    r = self.invoke_exact(cn, mn, md, cargs)
                elif op == 0xb7:
                    r = self.invoke_exact(cn, mn, md, cargs) if not isinstance(cargs[0], JNative) else self.native(cn, mn, md, cargs)
                else:
                    r = self.invoke_virtual(cn, mn, md, cargs)
                if rt != "V":
                    if rt == "J" and not isinstance(r, L):
                        r = L(r)
                    elif rt == "Z":
                        r = int(bool(r))
                    push(r)
                pc += 5 if op == 0xb9 else 3
            elif op == 0xba:                                # invokedynamic: only string concatenation (assert / log messages) is tolerated
                e = cp[u2(pc + 1)]
                mn, md = cf.nat(e[2])
                at, _ = parse_desc(md)
                indy_args = st[len(st) - len(at):] if at else []
                del st[len(st) - len(at):]
                # string concatenation (assert / log messages) yields a placeholder; a lambda becomes an opaque object that nothing on the
                # interpreted paths ever invokes (static initialisers store them in fields)
                if mn == "makeConcatWithConstants":        # StringConcatFactory: recipe with \x01 per dynamic argument, \x02 per constant
                    ref, bargs = c.bootstrap[e[1]]
                    recipe = cf.utf(cp[bargs[0]][1])
                    consts = [cp[b_] for b_ in bargs[1:]]
                    out_, ai, ci = [], 0, 0
                    for ch in recipe:
                        if ch == "\x01":
                            out_.append(self.j_string(indy_args[ai])); ai += 1
                        elif ch == "\x02":
                            k_ = consts[ci]; ci += 1
                            out_.append(cf.utf(k_[1]) if k_[0] == "String" else str(k_[1]))
                        else:
                            out_.append(ch)
                    push("".join(out_))
                else:
                    # LambdaMetafactory: bootstrap argument 1 is the handle of the synthetic method that implements the lambda
                    ref, bargs = c.bootstrap[e[1]]
                    mh = cp[bargs[1]]
                    me = cp[mh[2]]
                    tn, td = cf.nat(me[2])
                    lam = JNative("lambda", (cf.utf(cp[me[1]][1]), tn, td, mh[1]))
                    lam.f["captured"] = indy_args
                    push(lam)
                pc += 5
            elif op == 0xbb:
                cn = cf.utf(cp[u2(pc + 1)][1])
                k = self.load(cn)
                push(self.new_object(k) if k is not None else JNative(cn)); pc += 3
            elif op == 0xc0:                                # checkcast: trusted
                pc += 3
            elif op == 0xc1:
                v = pop()
                cn = cf.utf(cp[u2(pc + 1)][1])
                push(int(isinstance(v, JObj) and v.cls.is_subclass_of(cn))); pc += 3
            elif op in (0xac, 0xad, 0xae, 0xaf, 0xb0):
                self.n_insn += ninsn
                return pop()
            elif op == 0xb1:
                self.n_insn += ninsn
                return None
            elif op == 0xbf:
                ex = pop()
                raise JavaThrow(ex.cls.name if isinstance(ex, JObj) else ex.name, str(getattr(ex, "v", "")))
            elif op == 0xaa:                                # tableswitch
                v = pop()
                p0 = (pc + 4) & ~3
                dflt, lo, hi = struct.unpack_from(">iii", bc, p0)
                pc += struct.unpack_from(">i", bc, p0 + 12 + 4 * (v - lo))[0] if lo <= v <= hi else dflt
            elif op == 0xab:                                # lookupswitch
                v = pop()
                p0 = (pc + 4) & ~3
                dflt, npairs = struct.unpack_from(">ii", bc, p0)
                tgt = dflt
                for k in range(npairs):
                    m, off = struct.unpack_from(">ii", bc, p0 + 8 + 8 * k)
                    if m == v:
                        tgt = off
                        break
                pc += tgt
            elif op == 0xc2 or op == 0xc3:
                pop(); pc += 1
            else:
                raise NotImplementedError("opcode 0x%02x (%s) in %s.%s" % (op, jdis.OPS.get(op, ("?",))[0], c.name, key))
