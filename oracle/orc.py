"""ctypes binding of the CPU oracle (liborc.so).  TEST INFRASTRUCTURE ONLY — see slr_oracle.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC_RESULT = np.dtype([("bc", "<u8"), ("ed", "<i4"), ("ed_second", "<i4"), ("offset", "i1"), ("n_ins", "i1"),
                      ("n_del", "i1"), ("n_sub", "i1"), ("rank", "<i4"), ("flags", "<u4")], align=True)
assert BC_RESULT.itemsize == 24 or BC_RESULT.itemsize == 32

COLLIDE_RESULT = np.dtype([("bc", "<u8", (2,)), ("valid", "u1"), ("n_sub", "u1", (2,)), ("n_ins", "u1", (2,)), ("n_del", "u1", (2,)),
                           ("pad", "u1")], align=True)
assert COLLIDE_RESULT.itemsize == 24

F_ASSIGNED, F_EXCEPTION, F_TIE_UNPIN = 1, 2, 4
INT_MAX = 2147483647


class Match(C.Structure):
    _fields_ = [("read_seq", C.c_uint64), ("bc", C.c_uint64), ("ed", C.c_int32), ("offset", C.c_int32),
                ("n_sub", C.c_int32), ("n_ins", C.c_int32), ("n_del", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = [os.path.join(_HERE, f) for f in ("slr_oracle.c", "slr_oracle_assign.c", "slr_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_pack2bit.restype = C.c_uint64
        L.orc_pack2bit.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_revcomp2bit.restype = C.c_uint64
        L.orc_revcomp2bit.argtypes = [C.c_uint64, C.c_int]
        L.orc_replace_deg.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
        L.orc_insert_deg.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
        L.orc_delete_byte.restype = C.c_uint64
        L.orc_delete_byte.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.orc_set_new.restype = C.c_void_p
        L.orc_set_new.argtypes = [C.c_void_p, C.c_int64]
        L.orc_set_free.argtypes = [C.c_void_p]
        L.orc_set_find.restype = C.c_int64
        L.orc_set_find.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_match_tester.restype = C.c_int
        L.orc_match_tester.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.POINTER(Match), C.POINTER(C.c_int64)]
        L.orc_assign_barcode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_assign_barcode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                               C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.POINTER(C.c_int64), C.c_int]
        L.orc_exact_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_void_p]
        L.orc_collide_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        L.orc_limited_compare.restype = C.c_int
        L.orc_limited_compare.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.orc_umi_best9.restype = C.c_int32
        L.orc_umi_best9.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_umi_transpose.restype = C.c_int32
        L.orc_umi_transpose.argtypes = [C.c_int32]
        L.orc_umi_equality.restype = C.c_int32
        L.orc_umi_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.orc_umi_matrix_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_int]
        _LIB = L
    return _LIB


class BarcodeSet:
    """Search set (whitelist / used-barcode list) + ranks."""

    def __init__(self, keys, rank=None):
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self.rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
        self.h = lib().orc_set_new(self.keys.ctypes.data, len(self.keys))

    def __del__(self):
        try:
            lib().orc_set_free(self.h)
        except Exception:
            pass

    def find(self, key):
        return lib().orc_set_find(self.h, int(key))


def match_tester(bset, seq, length, ed, skip_full=False, allow_indels=True, post4=None, do_next=True, offset=0):
    out = (Match * 16)()
    probes = C.c_int64(0)
    if post4 is None:
        pp, pl = None, -1
    else:
        arr = np.ascontiguousarray(post4, dtype=np.uint8)
        pp, pl = arr.ctypes.data, len(arr)
    n = lib().orc_match_tester(bset.h, int(seq), length, ed, int(skip_full), int(allow_indels), pp, pl, int(do_next),
                               offset, out, C.byref(probes))
    if n < 0:
        return None, probes.value
    return [dict(read_seq=m.read_seq, bc=m.bc, ed=m.ed, offset=m.offset, n_sub=m.n_sub, n_ins=m.n_ins, n_del=m.n_del)
            for m in out[:n]], probes.value


def assign_barcode_batch(bset, slices, anchor, ed_max, plusminus=2, three_prime=True, bc_len=16, slice_len=None,
                         n_threads=0):
    """slices: uint8 [n, stride]; anchor: int32 [n].  Returns (results[BC_RESULT], total_probes)."""
    slices = np.ascontiguousarray(slices, dtype=np.uint8)
    anchor = np.ascontiguousarray(anchor, dtype=np.int32)
    n, stride = slices.shape
    if slice_len is None:
        slice_len = stride
    out = np.zeros(n, dtype=BC_RESULT)
    assert out.itemsize == 32
    probes = C.c_int64(0)
    lib().orc_assign_barcode_batch(bset.h, None if bset.rank is None else bset.rank.ctypes.data, ed_max, plusminus,
                                   int(three_prime), bc_len, slices.ctypes.data, stride, slice_len, anchor.ctypes.data,
                                   n, out.ctypes.data, C.byref(probes), n_threads)
    return out, probes.value


def exact_lookup_batch(bset, slices, anchor, three_prime=True, bc_len=16, lens=None):
    """UsedCellBCListGenerator$Worker: exact lookup of the offset-0 window.  Returns BC_RESULT[n]."""
    slices = np.ascontiguousarray(slices, dtype=np.uint8)
    anchor = np.ascontiguousarray(anchor, dtype=np.int32)
    n, stride = slices.shape
    out = np.zeros(n, dtype=BC_RESULT)
    lp = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
    lib().orc_exact_lookup_batch(bset.h, None if bset.rank is None else bset.rank.ctypes.data, int(three_prime), bc_len,
                                 slices.ctypes.data, stride, min(stride, 32), None if lp is None else lp.ctypes.data,
                                 anchor.ctypes.data, n, out.ctypes.data)
    return out


def collide_batch(bset, queries, ed, bc_len=16, n_threads=0):
    """BarcodeDatasetColissionTester: Matches of every query barcode against bset.  Returns (COLLIDE_RESULT[n], probes)."""
    q = np.ascontiguousarray(queries, dtype=np.uint64)
    out = np.zeros(len(q), dtype=COLLIDE_RESULT)
    probes = C.c_int64(0)
    lib().orc_collide_batch(bset.h, ed, bc_len, q.ctypes.data, len(q), out.ctypes.data, C.byref(probes), n_threads)
    return out, probes.value


def umi_matrix_batch(umis, job_offsets, umi_len=12, n_threads=0):
    """umis: uint8 [m, umi_len+2] 4-bit codes; job_offsets: int64 [n_jobs+1].  Returns (flat int32, out_offsets)."""
    umis = np.ascontiguousarray(umis, dtype=np.uint8)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    sizes = np.diff(job_offsets)
    out_offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes * sizes, out=out_offsets[1:])
    out = np.zeros(int(out_offsets[-1]), dtype=np.int32)
    lib().orc_umi_matrix_batch(umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(sizes),
                               out.ctypes.data, out_offsets.ctypes.data, n_threads)
    return out, out_offsets


CLUSTER_REC = np.dtype([("n_neighbours", "<i4"), ("best_key", "<i4"), ("best_count", "<i4"), ("n_ties", "<i4")])


def umi_cluster_batch(matrices, job_offsets, out_offsets, ed, member=None, rank=None, n_threads=0):
    """orc_umi_cluster_batch: ClusterOne_MyClustering.clusterLocal's neighbour counts and chosen entries, all jobs."""
    matrices = np.ascontiguousarray(matrices, dtype=np.int32)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    out_offsets = np.ascontiguousarray(out_offsets, dtype=np.int64)
    m = int(job_offsets[-1]) if len(job_offsets) else 0
    rec = np.zeros(m, dtype=CLUSTER_REC)
    if member is not None:
        member = np.ascontiguousarray(member, dtype=np.uint8)
    if rank is not None:
        rank = np.ascontiguousarray(rank, dtype=np.int32)
    L = lib()
    L.orc_umi_cluster_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_umi_cluster_batch.restype = None
    L.orc_umi_cluster_batch(matrices.ctypes.data, job_offsets.ctypes.data, out_offsets.ctypes.data, len(job_offsets) - 1, int(ed),
                            member.ctypes.data if member is not None else None, rank.ctypes.data if rank is not None else None,
                            rec.ctypes.data, n_threads)
    return rec


# ---- clustering + UMI assignment of a job (ClusterOneHierarchical.call, slr_oracle_assign.c) ---------------------------------------
ASSIGN_REC = np.dtype([("center", "<i4"), ("u1", "i1"), ("u2", "i1"), ("pos2", "i1"), ("off_mean", "i1"), ("flags", "<u2"),
                       ("cluster_size", "<u2"), ("n_clusters", "<i4")], align=True)
assert ASSIGN_REC.itemsize == 16
UA_ASSIGNED, UA_SKIPPED, UA_TIE_UNPIN, UA_DEEP = 1, 2, 4, 8


class AssignParams(C.Structure):
    """config.xml:270-278 + UMIparameters defaults: complete-link ED 2, single-link ED 1, switch above 3000 reads with a neighbour,
    foldDepthBelowMaxDiscardForClustering 50, ClusterOneHierarchical up to 100 reads"""
    _fields_ = [("ed_complete", C.c_int32), ("ed_single", C.c_int32), ("single_threshold", C.c_int32), ("fold_depth", C.c_int32),
                ("max_hier", C.c_int32), ("deep", C.c_int32)]

    def __init__(self, ed_complete=2, ed_single=1, single_threshold=3000, fold_depth=50, max_hier=100, deep=1):
        super().__init__(ed_complete, ed_single, single_threshold, fold_depth, max_hier, deep)


def umi_assign_batch(matrices, job_offsets, out_offsets, params=None, job_qv01=None, n_threads=0):
    """orc_umi_assign_batch: ClusterOneHierarchical.call for every job of at most params.max_hier reads -> ASSIGN_REC per read."""
    matrices = np.ascontiguousarray(matrices, dtype=np.int32)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    out_offsets = np.ascontiguousarray(out_offsets, dtype=np.int64)
    params = params or AssignParams()
    m = int(job_offsets[-1]) if len(job_offsets) else 0
    rec = np.zeros(m, dtype=ASSIGN_REC)
    qv = None if job_qv01 is None else np.ascontiguousarray(job_qv01, dtype=np.uint8)
    L = lib()
    L.orc_umi_assign_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_umi_assign_batch.restype = None
    L.orc_umi_assign_batch(matrices.ctypes.data, job_offsets.ctypes.data, out_offsets.ctypes.data, len(job_offsets) - 1, C.byref(params),
                           None if qv is None else qv.ctypes.data, rec.ctypes.data, n_threads)
    return rec


def fu_set_ops(keys, victims):
    """orc_fu_set_ops: fastutil IntOpenHashSet filled with `keys`, removeAll(victims) -> (order before, order after)"""
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    victims = np.ascontiguousarray(victims, dtype=np.int32)
    before, after = np.zeros(len(keys) + 1, dtype=np.int32), np.zeros(len(keys) + 1, dtype=np.int32)
    L = lib()
    L.orc_fu_set_ops.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_fu_set_ops.restype = C.c_int
    n = L.orc_fu_set_ops(keys.ctypes.data, len(keys), victims.ctypes.data, len(victims), int(keys.max()) if len(keys) else 0, before.ctypes.data,
                         after.ctypes.data)
    return before[:len(keys)].tolist(), after[:n].tolist()


# ---- Illumina-guided search -------------------------------------------------------------------------
GUIDED_HIT = np.dtype([("seq", "<u8"), ("n_sub", "i1"), ("n_ins", "i1"), ("n_del", "i1"), ("offset", "i1"), ("where", "u1"),
                       ("level", "u1"), ("pad", "<u2")], align=True)
assert GUIDED_HIT.itemsize == 16
GUIDED_RESULT = np.dtype([("seq", "<u8", (2,)), ("n_sub", "i1", (2,)), ("n_ins", "i1", (2,)), ("n_del", "i1", (2,)),
                          ("offset", "i1", (2,)), ("where", "u1", (2,)), ("n_distinct", "u1"), ("flags", "u1"), ("n_raw", "<i4"),
                          ("min_err_gene", "<i4"), ("pad", "<i4")], align=True)
assert GUIDED_RESULT.itemsize == 40
W_GENE, W_ALL, W_EMPTY, G_EXCEPTION = 1, 2, 4, 1


def guided_batch(group_keys, group_offsets, slices, anchor, group_id, ed, length, plusminus, post_len, bailout=-1,
                 bc_flavour=False, all_keys=None, all_ed=0, empty_keys=None, empty_ed=0, slice_len=None, raw_cap=0, n_threads=0):
    """orc_guided_batch.  Returns (GUIDED_RESULT[n], raw GUIDED_HIT[n, raw_cap] or None, probes)."""
    L = lib()
    L.orc_guided_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int]
    gk = np.ascontiguousarray(group_keys, dtype=np.uint64)
    go = np.ascontiguousarray(group_offsets, dtype=np.int64)
    slices = np.ascontiguousarray(slices, dtype=np.uint8)
    anchor = np.ascontiguousarray(anchor, dtype=np.int32)
    gid = np.ascontiguousarray(group_id, dtype=np.int32)
    n, stride = slices.shape
    edv = np.ascontiguousarray(np.broadcast_to(np.asarray(ed, dtype=np.int32), (n,)))
    ak = None if all_keys is None else np.ascontiguousarray(all_keys, dtype=np.uint64)
    ek = None if empty_keys is None else np.ascontiguousarray(empty_keys, dtype=np.uint64)
    out = np.zeros(n, dtype=GUIDED_RESULT)
    raw = np.zeros((n, raw_cap), dtype=GUIDED_HIT) if raw_cap else None
    probes = C.c_int64(0)
    L.orc_guided_batch(gk.ctypes.data, go.ctypes.data, len(go) - 1, None if ak is None else ak.ctypes.data, 0 if ak is None else len(ak),
                       all_ed, None if ek is None else ek.ctypes.data, 0 if ek is None else len(ek), empty_ed, int(bc_flavour), length,
                       plusminus, bailout, post_len, slices.ctypes.data, stride, stride if slice_len is None else slice_len,
                       anchor.ctypes.data, gid.ctypes.data, edv.ctypes.data, n, out.ctypes.data,
                       None if raw is None else raw.ctypes.data, raw_cap, C.byref(probes), n_threads)
    return out, raw, probes.value


def guided_tester(group, seq, length, ed, post4, bailout=-1, offset=0, bc_flavour=False, all_set=None, all_ed=0, empty_set=None, empty_ed=0,
                  cap=4096):
    """orc_guided_tester: the raw matchingList of one BCUMIEDtesterBase run (one window).  group / all_set / empty_set: BarcodeSet or None."""
    class Sets(C.Structure):
        _fields_ = [("group", C.c_void_p), ("all", C.c_void_p), ("all_ed", C.c_int), ("empty", C.c_void_p), ("empty_ed", C.c_int),
                    ("bc_flavour", C.c_int)]
    L = lib()
    L.orc_guided_tester.restype = C.c_int64
    L.orc_guided_tester.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int64, C.POINTER(C.c_int64)]
    s = Sets(None if group is None else group.h, None if all_set is None else all_set.h, all_ed, None if empty_set is None else empty_set.h,
             empty_ed, int(bc_flavour))
    out = np.zeros(cap, dtype=GUIDED_HIT)
    p4 = np.ascontiguousarray(post4, dtype=np.uint8)
    probes = C.c_int64(0)
    n = L.orc_guided_tester(C.byref(s), int(seq), length, ed, 1, p4.ctypes.data, len(p4), bailout, offset, out.ctypes.data, cap, C.byref(probes))
    return (None if n < 0 else out[:min(n, cap)]), n
