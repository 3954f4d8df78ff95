"""Freezes known-answer vectors from the CPU oracle into tests/golden/*.npz.

The reference ships no tests or golden outputs for this path and cannot run here (JVM bytecode, no JDK), so these
vectors pin the ORACLE against drift; they were produced by oracle/slr_oracle.c at the commit that added them and
cross-checked against the independent restatement oracle/pyref.py.  Re-run only when a reference-reading bug is
fixed:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import orc  # noqa: E402
import workloads  # noqa: E402


def main():
    for name, seed, tp, ed, skew, dense in [("bc_3p_ed1", 101, True, 1, False, False), ("bc_3p_ed2", 102, True, 2, False, False),
                                            ("bc_3p_ed2_skew", 103, True, 2, True, False), ("bc_5p_ed2", 104, False, 2, False, False),
                                            ("bc_3p_ed2_dense", 105, True, 2, True, True), ("bc_5p_ed1_dense", 106, False, 1, False, True),
                                            ("bc_3p_ed0", 107, True, 0, False, False)]:
        _, slices, anchors, wl = workloads.adversarial(seed, tp, 160, skew=skew, dense=dense)
        rank = np.arange(1, len(wl) + 1, dtype=np.int32)
        res, probes = orc.assign_barcode_batch(orc.BarcodeSet(wl, rank), slices, anchors, ed, 2, tp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), slices=slices, anchor=anchors, whitelist=wl, rank=rank,
                            ed=np.int32(ed), three_prime=np.int32(tp), result=res, probes=np.int64(probes))
        print(name, len(wl), int((res["flags"] & 1).sum()), probes)
    for name, seed, ed, skew in [("collide_ed1", 301, 1, False), ("collide_ed2", 302, 2, False), ("collide_ed2_skew", 303, 2, True)]:
        wl = workloads.used_list(seed, 150, skew)
        res, probes = orc.collide_batch(orc.BarcodeSet(wl), wl, ed)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), whitelist=wl, ed=np.int32(ed), result=res, probes=np.int64(probes))
        print(name, len(wl), int((res["valid"] & 1).sum()), int(((res["valid"] >> 1) & 1).sum()), probes)
    for umi_len in (12, 10):
        umis, offs = workloads.umi_jobs(200 + umi_len, umi_len, n_jobs=30, max_n=24)
        m, oo = orc.umi_matrix_batch(umis, offs, umi_len)
        np.savez_compressed(os.path.join(HERE, "umi_len%d.npz" % umi_len), umis=umis, job_offsets=offs, umi_len=np.int32(umi_len),
                            matrix=m, out_offsets=oo)
        print("umi", umi_len, len(m))
    # Illumina-guided search (a15): cross-checked against oracle/pyref.py by tests/test_guided.py::test_oracle_vs_python_restatement
    for name, seed, L, ed, pm, post_len, bailout, bc, nq in [("guided_umi_ed1", 401, 12, 1, 2, 5, -1, 0, 200), ("guided_umi_ed2", 402, 12, 2, 2, 6, -1, 0, 160),
                                                             ("guided_umi_ed2_bail1", 403, 12, 2, 1, 5, 1, 0, 160), ("guided_bc_ed2_bail2", 404, 16, 2, 2, 10, 2, 1, 120),
                                                             ("guided_umi_ed3", 405, 10, 3, 1, 6, -1, 0, 16), ("guided_bc_mixed", 406, 16, None, 2, 10, 2, 1, 160)]:
        w = workloads.guided(seed, L, nq, 2 if ed is None else ed, pm, post_len, bool(bc), skew=(seed % 2 == 0))
        edv = (np.random.default_rng(seed).integers(0, 3, nq) if ed is None else np.full(nq, ed)).astype(np.int32)
        res, raw, probes = orc.guided_batch(w["group_keys"], w["group_offsets"], w["slices"], w["anchor"], w["group_id"], edv, L, pm, post_len,
                                            bailout=bailout, bc_flavour=bool(bc), all_keys=w["all_keys"], all_ed=3, empty_keys=w["empty_keys"],
                                            empty_ed=2, slice_len=w["slice_len"], raw_cap=32)
        raw[res["flags"] != 0] = 0
        empty = np.zeros(0, dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), group_keys=w["group_keys"], group_offsets=w["group_offsets"],
                            all_keys=empty if w["all_keys"] is None else w["all_keys"], empty_keys=empty if w["empty_keys"] is None else w["empty_keys"],
                            slices=w["slices"], anchor=w["anchor"], group_id=w["group_id"], slice_len=np.int32(w["slice_len"]), ed=edv,
                            L=np.int32(L), pm=np.int32(pm), post_len=np.int32(post_len), bailout=np.int32(bailout), bc=np.int32(bc),
                            result=res, raw=raw, probes=np.int64(probes))
        print(name, nq, int((res["n_raw"] > 0).sum()), int((res["flags"] != 0).sum()), probes)


if __name__ == "__main__":
    main()
