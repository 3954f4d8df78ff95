"""Wider bytecode pin of the read-name format: FastqRecordExt.getScanDatFromReadName and getRecordForWriting run by the interpreter
on 1 000 + 1 000 more cases with other seeds (generators and layouts of oracle/make_ref_vectors.py).

    python oracle/make_ref_names_wide.py      -> tests/golden/ref_read_names_wide.npz, ref_written_names_wide.npz
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
    rn = M.read_name_cases(vm, np.random.default_rng(7131), 500)
    np.savez_compressed(os.path.join(M.OUT, "ref_read_names_wide.npz"), name=np.array([r[0] for r in rn]), limit=np.array([r[1] for r in rn], dtype=np.int32),
                        parsed=np.array([r[2] for r in rn]))
    print("getScanDatFromReadName", len(rn), "names")
    wn = M.write_name_cases(vm, np.random.default_rng(7152), 1000)
    np.savez_compressed(os.path.join(M.OUT, "ref_written_names_wide.npz"), name=np.array([c["name"] for c in wn]), stranded=np.array([c["stranded"] for c in wn]),
                        quals=np.array([c["quals"] for c in wn]), rev=np.array([c["rev"] for c in wn], dtype=np.int32),
                        five=np.array([c["five"] for c in wn], dtype=np.int32), read_id=np.array([c["rid"] for c in wn], dtype=np.int64),
                        kw=np.array([c["kw"] for c in wn]))
    print("getRecordForWriting", len(wn), "names")


if __name__ == "__main__":
    main()
