"""Golden vectors of the caller-side split of an oversized (cell, region) group FROM THE REFERENCE'S OWN CLASS FILES: UmiClustering.lambda$cluster$7
(F!com/rw/umifinder/analyzers/clustering/UmiClustering.class, UmiClustering.java:L136-L142) with commons-collections4's ListUtils.partition as
bytecode; MAX_SQUARE_NRECORDSPROCESSING = RAM_RESERVED / 300 (UmiClustering.java:L59) is injected.  Frozen in tests/golden/ref_split.npz.

    python oracle/make_ref_split.py
"""
import sys, glob, math, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import minijvm as J
REF="/root/reference/Jar"
jars=[REF+"/NanoporeBC_UMI_finder-2.1.jar"]+glob.glob(REF+"/lib/commons-collections4*.jar")
class SVM(J.VM):
    def native(self, cls, name, desc, args):
        a=args
        if cls=="java/lang/Math":
            if name=="sqrt": return J.D(math.sqrt(float(a[0])))
            if name=="ceil": return J.D(math.ceil(float(a[0])))
        if cls.endswith("Logger"): return None
        if name=="subList" and isinstance(a[0], J.JNative): return J.JNative("java/util/ArrayList", a[0].v[a[1]:a[2]])
        return super().native(cls,name,desc,args)
vm=SVM(jars)
c=vm.load("com/rw/umifinder/analyzers/clustering/UmiClustering")
c.initialized=True
class L:  # logger stub object
    pass
rows = []
for ram in (8e9, 16e9, 64e9, 128e9):
    c.statics["MAX_SQUARE_NRECORDSPROCESSING"]=int(int(ram)//300)
    c.statics["LOGGER"]=J.JNative("org/apache/logging/log4j/Logger", None)
    for n in (2, 100, 101, 5000, 7302, 7303, 7304, 10000, 14000, 14605, 14606, 14607, 14608, 20000, 21909, 21910, 29214, 29215, 50000, 100000):
        lst=J.JNative("java/util/ArrayList", list(range(n)))
        r=vm.run(c, "lambda$cluster$7(Ljava/util/List;)Ljava/util/List;", [lst])
        sz=vm.invoke_virtual(r.cls.name, "size", "()I", [r])
        parts=[len(vm.invoke_virtual(r.cls.name, "get", "(I)Ljava/lang/Object;", [r,i]).v) for i in range(sz)]
        print(int(ram / 1e9), n, parts)
        rows.append((int(ram), n, parts))

np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_split.npz"),
                    ram=np.array([r[0] for r in rows], dtype=np.int64), n=np.array([r[1] for r in rows], dtype=np.int64),
                    n_parts=np.array([len(r[2]) for r in rows], dtype=np.int64), parts=np.concatenate([np.array(r[2], dtype=np.int64) for r in rows]))
