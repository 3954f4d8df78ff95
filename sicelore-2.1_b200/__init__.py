"""sicelore-2.1_b200 — host-side mirror (Python, ctypes) of the reference interface for the barcode / UMI
edit-distance hot path, on top of the C ABI of libsicelore_gpu.so (include/sicelore_gpu.h; host-side entry points: include/sicelore_host.h).

The directory name is not a Python identifier; `__graft_entry__.load_package()` imports it as `sicelore_b200`.

Names follow the reference (F! = Jar/NanoporeBC_UMI_finder-2.1.jar):
  BarcodesMapForBCfinding    F!com/rw/nanoporereadscanner/WorkerReadscanner$BarcodesMapForBCfinding  (search set + ranks)
  Parser.assign_barcodes     F!com/rw/nanoporereadscanner/analyzers/Parser.assignBarcode (Parser.java:L195-L315)
  generate_distance_matrices F!com/rw/clustering/ClusteringEditDistanceBase.generateDistanceMatrix (…java:L168-L259)
  BestEditDistance           F!com/rw/clustering/ClusteringEditDistanceBase$BestEditDistance (L382-L449)
  read_name_suffix           F!com/rw/nanoporereadscanner/readerwriter/FastqRecordExt (bc= ed= ed_sec= bcStart= bcEnd= rk=)

There is NO CPU fallback here: every compute call goes to the CUDA library and raises when the library or a device
is missing.  (The CPU oracle lives in oracle/ and is test infrastructure only.)
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_GPU = os.environ.get("SLR_LIB_GPU") or os.path.join(_HERE, "libsicelore_gpu.so")   # override: kernel experiments only
LIB_SYNTH = os.path.join(_HERE, "libslr_synth.so")
NVCC = os.environ.get("SLR_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++"            # $CXX in this image points at a gcc without libgomp

SLR_OK, SLR_E_INVALID, SLR_E_NODEVICE, SLR_E_CUDA, SLR_E_NOMEM, SLR_E_UNSUPPORTED, SLR_E_REFERENCE_THROWS = 0, -1, -2, -3, -4, -5, -6
F_ASSIGNED, F_EXCEPTION, F_TIE_UNPIN = 1, 2, 4
INT_MAX = 2147483647

BC_RESULT = np.dtype([("bc", "<u8"), ("ed", "<i4"), ("ed_second", "<i4"), ("offset", "i1"), ("n_ins", "i1"),
                      ("n_del", "i1"), ("n_sub", "i1"), ("rank", "<i4"), ("flags", "<u4")], align=True)
assert BC_RESULT.itemsize == 32
COLLIDE_RESULT = np.dtype([("bc", "<u8", (2,)), ("valid", "u1"), ("n_sub", "u1", (2,)), ("n_ins", "u1", (2,)), ("n_del", "u1", (2,)),
                           ("pad", "u1")], align=True)
assert COLLIDE_RESULT.itemsize == 24
GUIDED_RESULT = np.dtype([("seq", "<u8", (2,)), ("n_sub", "i1", (2,)), ("n_ins", "i1", (2,)), ("n_del", "i1", (2,)),
                          ("offset", "i1", (2,)), ("where", "u1", (2,)), ("n_distinct", "u1"), ("flags", "u1"), ("n_raw", "<i4"),
                          ("min_err_gene", "<i4"), ("pad", "<i4")], align=True)
assert GUIDED_RESULT.itemsize == 40
GUIDED_HIT = np.dtype([("seq", "<u8"), ("n_sub", "i1"), ("n_ins", "i1"), ("n_del", "i1"), ("offset", "i1"), ("where", "u1"),
                       ("level", "u1"), ("pad", "<u2")], align=True)
assert GUIDED_HIT.itemsize == 16

EXPORTS = ["slr_ctx_create", "slr_ctx_destroy", "slr_ctx_device", "slr_bc_table_create", "slr_bc_table_destroy",
           "slr_bc_table_size", "slr_bc_assign", "slr_bc_assign_dev", "slr_bc_counts_read", "slr_bc_counts_reset",
           "slr_bc_counts_device", "slr_bc_exact", "slr_bc_exact_dev", "slr_bc_collide", "slr_bc_collide_dev", "slr_umi_dist", "slr_umi_dist_dev", "slr_umi_cluster", "slr_umi_cluster_dev", "slr_umi_assign", "slr_umi_assign_dev", "slr_umi_assign_dev2", "slr_umi_assign_scratch_bytes", "slr_umi_assign_deep_job_bytes", "slr_umi_session_assign", "slr_umi_session_create", "slr_umi_session_cluster",
           "slr_umi_session_matrices", "slr_umi_session_cells", "slr_umi_session_reads", "slr_umi_session_jobs", "slr_umi_session_destroy", "slr_guided_sets_create", "slr_guided_sets_destroy", "slr_guided_match", "slr_guided_match_dev",
           "slr_dyn_max_ed", "slr_last_error", "slr_abi_version", "slr_launch_count",
           "slr_multi_create", "slr_multi_destroy", "slr_multi_n_devices", "slr_multi_ctx", "slr_multi_peer_access", "slr_multi_bc_table_create",
           "slr_multi_bc_table_destroy", "slr_multi_bc_table_replica", "slr_multi_bc_assign", "slr_multi_bc_exact", "slr_multi_bc_counts_read",
           "slr_multi_bc_counts_reset", "slr_multi_umi_dist", "slr_multi_umi_cluster", "slr_multi_umi_assign",
           "slr_grouper_create", "slr_grouper_destroy", "slr_grouper_next_region_id", "slr_grouper_group_sams", "slr_group_jobs", "slr_needleman_errors", "slr_guided_mismatch_diff", "slr_bc_used_filter_low_counts", "slr_bc_used_merge_collisions"]


class SiceloreGpuError(RuntimeError):
    """A libsicelore_gpu call failed (code = SLR_E_*).  The Java shim maps this to log + System.exit(1)."""

    def __init__(self, code, msg):
        super().__init__("libsicelore_gpu error %d: %s" % (code, msg))
        self.code = code


class SiceloreGpuMissing(RuntimeError):
    """The CUDA library has not been built: there is no other implementation to fall back to."""


# ---------------------------------------------------------------------------------------------- build
def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    """Compile libsicelore_gpu.so (nvcc, sm_100a only) and libslr_synth.so (g++) in-tree."""
    srcs = [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC))]
    inc = os.path.join(_HERE, "..", "include", "sicelore_gpu.h")
    inc_host = os.path.join(_HERE, "..", "include", "sicelore_host.h")
    cu = [os.path.join(_CSRC, f) for f in ("slr_api.cu", "bc_assign.cu", "bc_collide.cu", "umi_dist.cu", "umi_cluster.cu", "umi_assign.cu", "umi_assign_deep.cu",
                                           "guided_match.cu", "slr_multi.cu", "slr_group.cpp", "slr_needleman.cpp", "slr_usedlist.cpp")]
    if force or _stale(LIB_GPU, srcs + [inc, inc_host]):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "-shared", "-ccbin", CXX, "-o", LIB_GPU] + cu
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    syn = os.path.join(_CSRC, "synth.cpp")
    if force or _stale(LIB_SYNTH, [syn]):
        subprocess.check_call([CXX, "-O3", "-fPIC", "-shared", "-fopenmp", "-std=c++17", "-o", LIB_SYNTH, syn])
    return LIB_GPU, LIB_SYNTH


# ---------------------------------------------------------------------------------------------- library handles
_gpu = None
_syn = None


def gpu_lib():
    """The CUDA library.  Fails loudly when it has not been built."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(LIB_GPU):
            raise SiceloreGpuMissing("libsicelore_gpu.so is missing: run __graft_entry__.build() (needs nvcc); "
                                     "there is no CPU fallback")
        L = C.CDLL(LIB_GPU)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.slr_ctx_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.slr_ctx_destroy.argtypes = [vp]
        L.slr_ctx_destroy.restype = None
        L.slr_ctx_device.argtypes = [vp]
        L.slr_bc_table_create.argtypes = [vp, vp, vp, i64, i32, C.POINTER(vp)]
        L.slr_grouper_create.argtypes = [i32, i64, C.POINTER(vp)]
        L.slr_grouper_destroy.argtypes = [vp]
        L.slr_grouper_destroy.restype = None
        L.slr_grouper_next_region_id.argtypes = [vp]
        L.slr_grouper_next_region_id.restype = i64
        L.slr_grouper_group_sams.argtypes = [vp, vp, vp, vp, i64, i32, vp, C.POINTER(i64)]
        L.slr_group_jobs.argtypes = [vp, vp, vp, i64, i32, i64, vp, vp, C.POINTER(i64)]
        L.slr_needleman_errors.argtypes = [C.c_uint64, C.c_uint64, i32, vp, vp]
        L.slr_bc_used_filter_low_counts.argtypes = [vp, i64, i64, vp]
        L.slr_bc_used_merge_collisions.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]
        L.slr_guided_mismatch_diff.argtypes = [vp, i64, vp, i32, i32, vp, i32, vp, vp]
        L.slr_bc_table_destroy.argtypes = [vp]
        L.slr_bc_table_destroy.restype = None
        L.slr_bc_table_size.argtypes = [vp]
        L.slr_bc_table_size.restype = i64
        L.slr_bc_assign.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, i64, vp]
        L.slr_bc_assign_dev.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, i64, vp, vp]
        L.slr_bc_counts_read.argtypes = [vp, vp, vp]
        L.slr_bc_counts_reset.argtypes = [vp, vp]
        L.slr_bc_counts_device.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
        L.slr_bc_exact.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, i64, vp]
        L.slr_bc_exact_dev.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, i64, vp, vp]
        L.slr_bc_collide.argtypes = [vp, vp, i32, vp, i64, vp]
        L.slr_bc_collide_dev.argtypes = [vp, vp, i32, vp, i64, vp, vp]
        L.slr_umi_dist.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
        L.slr_umi_dist_dev.argtypes = [vp, vp, i32, i32, vp, i64, i64, vp, vp, i64, vp]
        L.slr_umi_cluster.argtypes = [vp, vp, i32, i32, vp, i64, i32, vp, vp, vp, vp, vp]
        L.slr_umi_cluster_dev.argtypes = [vp, vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp]
        L.slr_umi_assign.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp, vp, vp, vp]
        L.slr_umi_assign_dev.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp]
        L.slr_umi_assign_dev2.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, vp, vp]
        L.slr_umi_assign_scratch_bytes.argtypes = [i64]
        L.slr_umi_assign_scratch_bytes.restype = i64
        L.slr_umi_assign_deep_job_bytes.argtypes = [i64]
        L.slr_umi_assign_deep_job_bytes.restype = i64
        L.slr_umi_session_assign.argtypes = [vp, vp, vp, vp]
        L.slr_umi_session_create.argtypes = [vp, vp, i32, i32, vp, i64, C.POINTER(vp)]
        L.slr_umi_session_cluster.argtypes = [vp, i32, vp, vp, vp]
        L.slr_umi_session_matrices.argtypes = [vp, vp, i64]
        L.slr_umi_session_cells.argtypes = [vp]
        L.slr_umi_session_cells.restype = i64
        L.slr_umi_session_destroy.argtypes = [vp]
        L.slr_umi_session_destroy.restype = None
        L.slr_guided_sets_create.argtypes = [vp, vp, vp, i64, vp, i64, i32, vp, i64, i32, i32, i32, C.POINTER(vp)]
        L.slr_guided_sets_destroy.argtypes = [vp]
        L.slr_guided_sets_destroy.restype = None
        L.slr_guided_match.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, i64, vp, vp, i32]
        L.slr_guided_match_dev.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, i32, i64, vp, vp, i32, vp]
        L.slr_dyn_max_ed.argtypes = [vp, i32, i32, i32, i32]
        L.slr_multi_create.argtypes = [i32, vp, i32, C.POINTER(vp)]
        L.slr_multi_destroy.argtypes = [vp]
        L.slr_multi_destroy.restype = None
        L.slr_multi_n_devices.argtypes = [vp]
        L.slr_multi_ctx.argtypes = [vp, i32]
        L.slr_multi_ctx.restype = vp
        L.slr_multi_peer_access.argtypes = [vp]
        L.slr_multi_bc_table_create.argtypes = [vp, vp, vp, i64, i32, C.POINTER(vp)]
        L.slr_multi_bc_table_destroy.argtypes = [vp]
        L.slr_multi_bc_table_destroy.restype = None
        L.slr_multi_bc_table_replica.argtypes = [vp, i32]
        L.slr_multi_bc_table_replica.restype = vp
        L.slr_multi_bc_assign.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, i64, vp]
        L.slr_multi_bc_exact.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, i64, vp]
        L.slr_multi_bc_counts_read.argtypes = [vp, vp, vp]
        L.slr_multi_bc_counts_reset.argtypes = [vp, vp]
        L.slr_multi_umi_dist.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
        L.slr_multi_umi_cluster.argtypes = [vp, vp, i32, i32, vp, i64, i32, vp, vp, vp]
        L.slr_multi_umi_assign.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp, vp]
        L.slr_last_error.restype = C.c_char_p
        L.slr_abi_version.restype = i32
        L.slr_launch_count.restype = i64
        _gpu = L
    return _gpu


def _check(rc):
    if rc != 0:
        raise SiceloreGpuError(rc, gpu_lib().slr_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(gpu_lib().slr_launch_count())


def lib_sha256():
    """sha256 of the libsicelore_gpu.so that is loaded: ties a profile under profiles/ to the binary a bench line was measured with"""
    import hashlib
    with open(LIB_GPU, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def csrc_sha256():
    """sha256 over the CUDA / C++ sources of the library (csrc/ + the public header), in sorted file order: survives a rebuild"""
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f != "synth.cpp"] + \
            [os.path.join(_HERE, "..", "include", f) for f in ("sicelore_gpu.h", "sicelore_host.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


KERNEL_TU = {"bc_assign_kernel": "bc_assign.cu", "umi_pairs_kernel": "umi_dist.cu", "umi_assign_kernel": "umi_assign.cu",
             "guided_match_kernel": "guided_match.cu", "umi_assign_deep_kernel": "umi_assign_deep.cu"}


def kernel_src_sha256(kernel):
    """sha256 over the translation unit of `kernel` (name up to '<', '/', '@') and every header it includes, transitively, from csrc/ and include/:
    a profile entry describes a kernel as long as THESE files are unchanged, whatever happens to the other translation units"""
    import hashlib
    import re
    base = re.split(r"[<@/]", kernel)[0]
    if base not in KERNEL_TU:
        return None
    inc_dir = os.path.join(_HERE, "..", "include")
    seen, todo = {}, [os.path.join(_CSRC, KERNEL_TU[base])]
    while todo:
        f = todo.pop()
        name = os.path.basename(f)
        if name in seen:
            continue
        with open(f, "rb") as fh:
            data = fh.read()
        seen[name] = data
        for inc in re.findall(rb'#include\s+"([^"]+)"', data):
            for d in (_CSRC, inc_dir):
                cand = os.path.join(d, inc.decode())
                if os.path.exists(cand):
                    todo.append(cand)
                    break
    h = hashlib.sha256()
    for name in sorted(seen):
        h.update(name.encode())
        h.update(seen[name])
    return h.hexdigest()


class SynthParams(C.Structure):
    _fields_ = [("p_sub", C.c_double), ("p_ins", C.c_double), ("p_del", C.c_double), ("p_random", C.c_double),
                ("p_n", C.c_double), ("jitter", C.c_double * 5), ("n_cells", C.c_int64), ("three_prime", C.c_int)]


def synth_lib():
    global _syn
    if _syn is None:
        if not os.path.exists(LIB_SYNTH):
            build()
        L = C.CDLL(LIB_SYNTH)
        L.slr_synth_whitelist.argtypes = [C.c_void_p, C.c_int64, C.c_uint64]
        L.slr_synth_whitelist.restype = None
        L.slr_synth_default_params.argtypes = [C.POINTER(SynthParams)]
        L.slr_synth_default_params.restype = None
        L.slr_synth_reads.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.POINTER(SynthParams),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.slr_synth_reads.restype = None
        L.slr_synth_umi_jobs.argtypes = [C.c_int64, C.c_double, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_int,
                                         C.c_void_p, C.c_void_p]
        L.slr_synth_umi_jobs.restype = None
        L.slr_synth_umi_jobs_at.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_int,
                                            C.c_void_p, C.c_void_p]
        L.slr_synth_umi_jobs_at.restype = None
        _syn = L
    return _syn


# ---------------------------------------------------------------------------------------------- synthetic workloads
def synth_whitelist(n, seed):
    """n distinct pseudo-random 16-mers, 2-bit packed (stand-in for 737K-august-2016 / 3M-february-2018)."""
    out = np.empty(n, dtype=np.uint64)
    synth_lib().slr_synth_whitelist(out.ctypes.data, n, seed)
    return out


def synth_reads(whitelist, n, seed, first=0, three_prime=True, n_cells=0, p_sub=0.02, p_ins=0.01, p_del=0.02,
                p_random=0.10, p_n=0.01, out=None):
    """Boundary buffers for reads [first, first+n) of a synthetic run: (slices[n,32] u8, anchor[n] i32, truth[n] i64)."""
    prm = SynthParams()
    synth_lib().slr_synth_default_params(C.byref(prm))
    prm.p_sub, prm.p_ins, prm.p_del, prm.p_random, prm.p_n = p_sub, p_ins, p_del, p_random, p_n
    prm.n_cells, prm.three_prime = n_cells, int(three_prime)
    wl = np.ascontiguousarray(whitelist, dtype=np.uint64)
    if out is None:
        slices = np.empty((n, 32), dtype=np.uint8)
        anchor = np.empty(n, dtype=np.int32)
    else:
        slices, anchor = out
    truth = np.empty(n, dtype=np.int64)
    synth_lib().slr_synth_reads(wl.ctypes.data, len(wl), first, n, seed, C.byref(prm), slices.ctypes.data,
                                anchor.ctypes.data, truth.ctypes.data, None)
    return slices, anchor, truth


def synth_umi_jobs(n_jobs, mean=4.0, cap=2000, seed=4, p_err=0.05, p_shift=0.05, umi_len=12):
    """(umis[m,16] u8 4-bit codes, job_offsets[n_jobs+1] i64) for n_jobs (cell, region) groups."""
    offs = np.zeros(n_jobs + 1, dtype=np.int64)
    synth_lib().slr_synth_umi_jobs(n_jobs, mean, cap, seed, p_err, p_shift, umi_len, offs.ctypes.data, None)
    umis = np.zeros((int(offs[-1]), 16), dtype=np.uint8)
    synth_lib().slr_synth_umi_jobs(n_jobs, mean, cap, seed, p_err, p_shift, umi_len, offs.ctypes.data, umis.ctypes.data)
    return umis, offs


def synth_umi_job_sizes(first_job, n_jobs, mean=4.0, cap=2000, seed=4):
    """sizes of jobs first_job .. first_job + n_jobs - 1 of the global synthetic job stream (job j depends on (seed, j) only)"""
    offs = np.zeros(n_jobs + 1, dtype=np.int64)
    synth_lib().slr_synth_umi_jobs_at(first_job, n_jobs, mean, cap, seed, 0.05, 0.05, 12, offs.ctypes.data, None)
    return np.diff(offs)


def synth_umi_jobs_at(first_job, n_jobs, mean=4.0, cap=2000, seed=4, p_err=0.05, p_shift=0.05, umi_len=12):
    """(umis, job_offsets) of jobs first_job .. first_job + n_jobs - 1 of the global stream"""
    offs = np.zeros(n_jobs + 1, dtype=np.int64)
    synth_lib().slr_synth_umi_jobs_at(first_job, n_jobs, mean, cap, seed, p_err, p_shift, umi_len, offs.ctypes.data, None)
    umis = np.zeros((int(offs[-1]), 16), dtype=np.uint8)
    synth_lib().slr_synth_umi_jobs_at(first_job, n_jobs, mean, cap, seed, p_err, p_shift, umi_len, offs.ctypes.data, umis.ctypes.data)
    return umis, offs


def synth_umi_shard(read_first, n_reads, mean=4.0, cap=2000, seed=4):
    """Reads [read_first, read_first + n_reads) of the global (cell, region)-sorted synthetic read stream, the way a sorted BAM is cut when
    it is sharded by read index: the first and the last job of the shard may be pieces of a job that continues on the neighbouring shard.
    Returns (umis [n_reads, 16], job_offsets (local), first_job_id, last_job_id): the ids are the jobs' keys in the global stream
    (cell = id // 2000, region = id % 2000 in the 5 k cells x 2 k genes picture)."""
    # global job offsets up to the end of the shard: sizes in blocks until the shard end is covered
    block = max(1 << 16, int((read_first + n_reads) / mean * 1.02) + 1024)
    sizes = synth_umi_job_sizes(0, block, mean, cap, seed)
    csum = np.cumsum(sizes)
    while csum[-1] < read_first + n_reads:
        more = synth_umi_job_sizes(len(sizes), block, mean, cap, seed)
        sizes = np.concatenate([sizes, more])
        csum = np.cumsum(sizes)
    goff = np.concatenate([[0], csum])
    j_lo = int(np.searchsorted(goff, read_first, side="right")) - 1
    j_hi = int(np.searchsorted(goff, read_first + n_reads, side="left")) - 1          # last job with a read in the shard
    umis, offs = synth_umi_jobs_at(j_lo, j_hi - j_lo + 1, mean, cap, seed)
    a = read_first - int(goff[j_lo])
    umis = np.ascontiguousarray(umis[a:a + n_reads])
    loc = np.clip(offs - a, 0, n_reads)
    return umis, np.ascontiguousarray(loc, dtype=np.int64), j_lo, j_hi


# ---------------------------------------------------------------------------------------------- 2-bit helpers
_B2 = "AGCT"            # NucleicAcidTwoBitPerBase: A=0 G=1 C=2 T=3 (T!…NucleicAcidTwoBitPerBase.java:L80-L87)


def synth_guided(n, seq_len, seed=9, n_groups=1000, group_size=8, pm=2, post_len=8, bc_flavour=False, n_all=6000, n_empty=100000,
                 p_sub=0.02, p_ins=0.01, p_del=0.02, p_random=0.1):
    """Synthetic Illumina-guided batch (numpy, vectorised): n_groups candidate groups of ~group_size random sequences, every read is a
    candidate of its group pushed through the error model of synth_reads (sub / ins / del per base), placed at anchor = pm + 1 of a
    32-byte stranded slice with +-1 jitter of the predicted position; p_random of the reads carry a random window.
    Returns dict(group_keys, group_offsets, all_keys, empty_keys, slices, anchor, group_id)."""
    rng = np.random.default_rng(seed)
    L = int(seq_len)
    sizes = np.maximum(1, rng.poisson(group_size, n_groups)) if group_size > 1 else np.ones(n_groups, dtype=np.int64)
    go = np.zeros(n_groups + 1, dtype=np.int64)
    np.cumsum(sizes, out=go[1:])
    gk = rng.integers(0, 1 << (2 * L), int(go[-1]), dtype=np.uint64)
    gid = rng.integers(0, n_groups, n).astype(np.int32)
    pick = go[gid] + (rng.random(n) * sizes[gid]).astype(np.int64)
    true = gk[pick]
    rnd = rng.random(n) < p_random
    true = np.where(rnd, rng.integers(0, 1 << (2 * L), n, dtype=np.uint64), true)
    bases = np.frombuffer(b"AGCT", dtype=np.uint8)
    digits = ((true[:, None] >> (2 * (L - 1 - np.arange(L, dtype=np.uint64)))[None, :]) & np.uint64(3)).astype(np.int64)   # [n, L]
    tail = rng.integers(0, 4, (n, 40))
    src = np.concatenate([digits, tail], axis=1)                                   # candidate followed by random bases
    # per-base edits: walk the source with a cursor; del skips a source base, ins emits a random base without advancing
    out = np.zeros((n, 32), dtype=np.int64)
    lead = pm + 1 + rng.choice([-1, 0, 0, 0, 1], n) * (pm > 0)
    pre = rng.integers(0, 4, (n, 8))
    cur = np.zeros(n, dtype=np.int64)
    for col in range(32):
        in_lead = col < lead
        r = rng.random(n)
        is_ins = (~in_lead) & (r < p_ins)
        is_del = (~in_lead) & (r >= p_ins) & (r < p_ins + p_del)
        cur = cur + is_del
        b = src[np.arange(n), np.minimum(cur, src.shape[1] - 1)]
        sub = (~in_lead) & (rng.random(n) < p_sub)
        b = np.where(sub, (b + rng.integers(1, 4, n)) & 3, b)
        b = np.where(is_ins, rng.integers(0, 4, n), b)
        b = np.where(in_lead, pre[:, col % 8], b)
        out[:, col] = b
        cur = cur + ((~in_lead) & (~is_ins))
    slices = bases[out].astype(np.uint8)
    anchor = np.full(n, pm + 1, dtype=np.int32)
    ak = ek = None
    if bc_flavour:
        ak = np.unique(np.concatenate([gk[rng.integers(0, len(gk), min(n_all, len(gk)))], rng.integers(0, 1 << (2 * L), n_all // 4, dtype=np.uint64)]))
        ek = np.unique(rng.integers(0, 1 << (2 * L), n_empty, dtype=np.uint64))
    return dict(group_keys=gk, group_offsets=go, all_keys=ak, empty_keys=ek, slices=np.ascontiguousarray(slices), anchor=anchor, group_id=gid)


def pack_barcode(s):
    h = 0
    for ch in s:
        h = (h << 2) | _B2.index(ch.upper())
    return h


def unpack_barcode(h, length=16):
    return "".join(_B2[(int(h) >> (2 * (length - 1 - i))) & 3] for i in range(length))


def read_whitelist(path):
    """NanoporeReadScannerMain.readBarcodesFile: one barcode per line, optional '-1' suffix, .gz ok."""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    keys = []
    with op(path, "rt") as f:
        for line in f:
            s = line.strip().split("-")[0].split()[0] if line.strip() else ""
            if s:
                keys.append(pack_barcode(s))
    return np.array(keys, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------- context / table
class Context:
    """One per (process, device): streams and staging buffers of the CUDA library."""

    def __init__(self, device=-1, n_streams=2):
        h = C.c_void_p()
        _check(gpu_lib().slr_ctx_create(device, n_streams, C.byref(h)))
        self.h = h

    @property
    def device(self):
        return gpu_lib().slr_ctx_device(self.h)

    def close(self):
        if self.h:
            gpu_lib().slr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BarcodesMapForBCfinding:
    """Device-resident search set: barcode (2-bit long) -> CountsRank.rank, plus the per-barcode ED counters."""

    def __init__(self, ctx, barcodes2bit, rank=None, bc_len=16):
        self.ctx = ctx
        self.keys = np.ascontiguousarray(barcodes2bit, dtype=np.uint64)
        self.rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
        h = C.c_void_p()
        _check(gpu_lib().slr_bc_table_create(ctx.h, self.keys.ctypes.data, None if self.rank is None else self.rank.ctypes.data,
                                             len(self.keys), bc_len, C.byref(h)))
        self.h = h

    @classmethod
    def getMapFromCellRangerData(cls, ctx, barcodes2bit):
        """--cellRangerBCs flow: every barcode of the list gets a rank (1-based list order here)."""
        return cls(ctx, barcodes2bit, np.arange(1, len(barcodes2bit) + 1, dtype=np.int32))

    def size(self):
        return int(gpu_lib().slr_bc_table_size(self.h))

    def counts(self):
        """assignedBarcodes2ndPass as an [n, 3] array (reads assigned to barcode i at ED 0/1/2) -> BarcodesAssigned.tsv"""
        out = np.zeros((len(self.keys), 3), dtype=np.int64)
        _check(gpu_lib().slr_bc_counts_read(self.ctx.h, self.h, out.ctypes.data))
        return out

    def reset_counts(self):
        _check(gpu_lib().slr_bc_counts_reset(self.ctx.h, self.h))

    def counts_device_ptr(self):
        p, n = C.c_void_p(), C.c_int64()
        _check(gpu_lib().slr_bc_counts_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self):
        if getattr(self, "h", None):
            gpu_lib().slr_bc_table_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Parser:
    """Mirror of the second-pass Parser for the part that moved to the GPU (assignBarcode, Parser.java:L195-L315)."""

    def __init__(self, ctx, hashMapForBCfinding, bcEditDistance=1, testPlusMinusPos=2, three_prime=True):
        self.ctx, self.map = ctx, hashMapForBCfinding
        self.ed, self.pm, self.three_prime = int(bcEditDistance), int(testPlusMinusPos), bool(three_prime)

    def assign_barcodes(self, slices, anchor, lens=None, out=None):
        """slices u8 [n, stride] (host), anchor i32 [n] -> structured array BC_RESULT [n]."""
        slices = np.ascontiguousarray(slices, dtype=np.uint8)
        anchor = np.ascontiguousarray(anchor, dtype=np.int32)
        n, stride = slices.shape
        if out is None:
            out = np.empty(n, dtype=BC_RESULT)
        lp = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.int32)
            lp = lens.ctypes.data
        _check(gpu_lib().slr_bc_assign(self.ctx.h, self.map.h, self.ed, self.pm, int(self.three_prime), slices.ctypes.data,
                                       stride, min(stride, 32), lp, anchor.ctypes.data, n, out.ctypes.data))
        return out

    def assign_barcodes_dev(self, d_slices, stride, d_anchor, n, d_out, stream=0, d_lens=0, slice_len=32):
        """Device-pointer variant (ints = raw CUDA pointers, e.g. torch tensor.data_ptr()); asynchronous on `stream`."""
        _check(gpu_lib().slr_bc_assign_dev(self.ctx.h, self.map.h, self.ed, self.pm, int(self.three_prime), d_slices, stride,
                                           slice_len, d_lens or None, d_anchor, n, d_out, stream or None))

    @staticmethod
    def barcode_positions(res, adapterpos, three_prime=True, bc_len=16):
        """BarcodeResult.start / end as the Java sets them (Parser.java:L273-L280); adapterpos = 1-based adapter end."""
        off = res["offset"].astype(np.int64)
        d = res["n_ins"].astype(np.int64) - res["n_del"].astype(np.int64)      # OneMatch.getOffsetForReadEnd
        if three_prime:
            start = adapterpos - 1 + off
            end = start - (bc_len - 1) - d
        else:
            start = adapterpos + 1 + off
            end = start + (bc_len - 1) + d
        return start, end


class UsedCellBCListGenerator:
    """Mirror of the pass-1 used-barcode counting (UsedCellBCListGenerator$Worker.call, UsedCellBCListGenerator.java:L206-L232):
    exact lookup of the offset-0 window in the 10x whitelist; counts accumulate in the table (unfilteredUsedBarcodeMap)."""

    def __init__(self, ctx, whitelist_table, three_prime=True):
        self.ctx, self.map, self.three_prime = ctx, whitelist_table, bool(three_prime)

    def addFastqs(self, slices, anchor, lens=None, out=None):
        slices = np.ascontiguousarray(slices, dtype=np.uint8)
        anchor = np.ascontiguousarray(anchor, dtype=np.int32)
        n, stride = slices.shape
        if out is None:
            out = np.empty(n, dtype=BC_RESULT)
        lp = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.int32)
            lp = lens.ctypes.data
        _check(gpu_lib().slr_bc_exact(self.ctx.h, self.map.h, int(self.three_prime), slices.ctypes.data, stride, min(stride, 32), lp,
                                      anchor.ctypes.data, n, out.ctypes.data))
        return out

    def unfilteredUsedBarcodeMap(self):
        """barcode -> read count for every whitelist barcode seen at least once (key order of the whitelist)"""
        c = self.map.counts()[:, 0]
        sel = np.nonzero(c)[0]
        return self.map.keys[sel], c[sel]


class BarcodeDatasetColissionTester:
    """Mirror of the pass-1 collision tester for the part that moved to the GPU: one BarcodeMatchTester run per used barcode
    against the list itself (BarcodeDatasetColissionTester.submitSeq, BarcodeDatasetColissionTester.java:L212-L229)."""

    def __init__(self, ctx, barcodes_b4filtering, editDistance):
        self.ctx, self.map, self.ed = ctx, barcodes_b4filtering, int(editDistance)

    def colissionsFromScan(self, barcodes=None, out=None):
        """Matches per barcode as a COLLIDE_RESULT array (default: every key of the map, in key order of the input)."""
        q = np.ascontiguousarray(self.map.keys if barcodes is None else barcodes, dtype=np.uint64)
        if out is None:
            out = np.empty(len(q), dtype=COLLIDE_RESULT)
        _check(gpu_lib().slr_bc_collide(self.ctx.h, self.map.h, self.ed, q.ctypes.data, len(q), out.ctypes.data))
        return out


UL_ORDER_UNPIN, UL_RANK_TIES = 1, 2


def used_filter_low_counts(counts, record_count):
    """UsedBarcodesListData.filterLowCounts with finalizeData's cutoff (slr_bc_used_filter_low_counts): bool mask of the barcodes that go on to
    the collision test"""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    keep = np.zeros(len(counts), dtype=np.uint8)
    _check(gpu_lib().slr_bc_used_filter_low_counts(counts.ctypes.data, len(counts), int(record_count), keep.ctypes.data))
    return keep.astype(bool)


def used_merge_collisions(barcodes, counts, collide, min_count_fold=10, merge_ed=1, cells_fold=500):
    """BarcodeDatasetColissionTester.generateColissionMergedBCmap + the ranks of WorkerReadscanner.java:L264-L270 on the records of
    BarcodeDatasetColissionTester.colissionsFromScan (slr_bc_used_merge_collisions).  Returns (keep mask, rank, flags): the kept barcodes with
    their ranks are the BarcodesMapForBCfinding of pass 2."""
    barcodes = np.ascontiguousarray(barcodes, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    collide = np.ascontiguousarray(collide, dtype=COLLIDE_RESULT)
    n = len(barcodes)
    assert counts.shape == (n,) and collide.shape == (n,)
    keep, rank, flags = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.int32), C.c_uint32(0)
    _check(gpu_lib().slr_bc_used_merge_collisions(barcodes.ctypes.data, counts.ctypes.data, collide.ctypes.data, n, int(min_count_fold), int(merge_ed),
                                                  int(cells_fold), keep.ctypes.data, rank.ctypes.data, C.byref(flags)))
    return keep.astype(bool), rank, int(flags.value)


def read_name_suffix(res_i, bc_start, bc_end):
    """The barcode part of FastqRecordExt's read-name suffix (README.md:400): bc= ed= ed_sec= bcStart= bcEnd= rk="""
    return "bc=%s_ed=%d_ed_sec=%d_bcStart=%d_bcEnd=%d_rk=%d" % (unpack_barcode(res_i["bc"]), res_i["ed"], res_i["ed_second"],
                                                              bc_start, bc_end, res_i["rank"])



# ---------------------------------------------------------------------------------------------- Illumina-guided search
class DynamicEditDistances:
    """Mirror of com.rw.parameters.DynamicEditDistances for one (length, errorpercent) column of bcMaxEditDistances.xml /
    umiMaxEditDistances.xml: max_candidates[e] = <maxBarcodes> of edit distance e (DynamicEditDistances.java:L75, L93-L98)."""

    def __init__(self, max_candidates):
        self.col = np.ascontiguousarray(max_candidates, dtype=np.int64)

    @classmethod
    def from_xml(cls, path, length, errorpercent):
        """Reads the reference's own table file (XMLData -> OneUMILengthColumn -> ErrorPercentColumn -> EditDistanceColumn)."""
        import xml.etree.ElementTree as ET
        for lc in ET.parse(path).getroot().iter():
            if lc.find("umiBCLength") is not None and int(lc.find("umiBCLength").text) == length:
                for ec in lc.iter():
                    if ec.find("errorpercent") is not None and int(ec.find("errorpercent").text) == errorpercent:
                        rows = {int(d.find("editDistance").text): int(d.find("maxBarcodes").text)
                                for d in ec.iter() if d.find("editDistance") is not None and d.find("maxBarcodes") is not None}
                        return cls([rows[e] for e in range(len(rows))])
        raise KeyError("no column for length %d, error %d %% in %s" % (length, errorpercent, path))

    def getmaxED(self, count, pos_plusminus, maxED=None):
        r = gpu_lib().slr_dyn_max_ed(self.col.ctypes.data, len(self.col), int(count), int(pos_plusminus), -1 if maxED is None else int(maxED))
        if r < 0:
            raise LookupError("NoSuchElementException: no edit distance admits %d candidates" % count)
        return r


class GuidedSets:
    """Device-resident candidate sets of the Illumina-guided search and the batched search itself: the offset loops of
    IlluminaUMIanalyzer.findUMI (IlluminaUMIanalyzer.java:L89-L136, bc_flavour=False: UMInucTwoBitPerBaseEDtester) and
    IlluminaBarcodeAnalyzer.testBarcodes (IlluminaBarcodeAnalyzer.java:L272-L304, bc_flavour=True: BCnucTwoBitPerBaseEDtester)
    plus the sorted().distinct() reduction of getBestAndSecondBCorUMI."""
    W_GENE, W_ALL, W_EMPTY, EXCEPTION = 1, 2, 4, 1

    def __init__(self, ctx, group_keys, group_offsets, seq_len, bc_flavour=False, all_keys=None, all_ed=0, empty_keys=None, empty_ed=0):
        self.ctx, self.seq_len = ctx, int(seq_len)
        gk = np.ascontiguousarray(group_keys, dtype=np.uint64)
        go = np.ascontiguousarray(group_offsets, dtype=np.int64)
        ak = None if all_keys is None else np.ascontiguousarray(all_keys, dtype=np.uint64)
        ek = None if empty_keys is None else np.ascontiguousarray(empty_keys, dtype=np.uint64)
        h = C.c_void_p()
        _check(gpu_lib().slr_guided_sets_create(ctx.h, gk.ctypes.data, go.ctypes.data, len(go) - 1, None if ak is None else ak.ctypes.data,
                                                0 if ak is None else len(ak), int(all_ed), None if ek is None else ek.ctypes.data,
                                                0 if ek is None else len(ek), int(empty_ed), int(bool(bc_flavour)), self.seq_len, C.byref(h)))
        self.h = h

    def match(self, slices, anchor, group_id, ed, posplusminus, post_len, bailout=None, slice_len=None, raw_cap=0):
        """slices uint8 [n, stride] (stranded orientation), anchor / group_id int32 [n], ed scalar or int32 [n].
        Returns (GUIDED_RESULT[n], GUIDED_HIT[n, raw_cap] or None)."""
        slices = np.ascontiguousarray(slices, dtype=np.uint8)
        anchor = np.ascontiguousarray(anchor, dtype=np.int32)
        gid = np.ascontiguousarray(group_id, dtype=np.int32)
        n, stride = slices.shape
        edv = np.ascontiguousarray(np.broadcast_to(np.asarray(ed, dtype=np.int32), (n,)))
        out = np.empty(n, dtype=GUIDED_RESULT)
        raw = np.empty((n, raw_cap), dtype=GUIDED_HIT) if raw_cap else None
        _check(gpu_lib().slr_guided_match(self.ctx.h, self.h, int(posplusminus), int(post_len), -1 if bailout is None else int(bailout),
                                          slices.ctypes.data, stride, stride if slice_len is None else int(slice_len), anchor.ctypes.data,
                                          gid.ctypes.data, edv.ctypes.data, n, out.ctypes.data, None if raw is None else raw.ctypes.data,
                                          int(raw_cap)))
        return out, raw

    def close(self):
        if getattr(self, "h", None):
            gpu_lib().slr_guided_sets_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

G_NO_SECOND = -(1 << 31)


class NeedlemanScores(C.Structure):
    """NeedlemanScores (T!com/rw/nuc/alignment/needleman/NeedlemanScores.class, …java:L44-L56), defaults as in the reference"""
    _fields_ = [(k, C.c_int32) for k in ("leading_gap_1", "leading_gap_2", "trailing_gap_1", "trailing_gap_2", "indel", "mismatch", "match")]

    def __init__(self, leading_gap_1=-4, leading_gap_2=-5, trailing_gap_1=-5, trailing_gap_2=-5, indel=-5, mismatch=-5, match=5):
        super().__init__(leading_gap_1, leading_gap_2, trailing_gap_1, trailing_gap_2, indel, mismatch, match)


def needleman_errors(template2bit, read2bit, length, scores=None):
    """NeedlemanWunsch(template, read, scores) + NeedlemanMatch.countNeedlemanErrorsInRead (slr_needleman_errors): (insertions, deletions,
    substitutions, total) of the alignment of two 2-bit packed sequences"""
    cnt = (C.c_int32 * 4)()
    _check(gpu_lib().slr_needleman_errors(int(template2bit), int(read2bit), int(length), C.byref(scores) if scores is not None else None, cnt))
    return tuple(cnt)


def guided_mismatch_diff(res, slices, anchor, seq_len, slice_len=None, scores=None):
    """nMismatchDiffBestvsSecondBest per record of GuidedSets.match (slr_guided_mismatch_diff; IlluminaBarcodeUMIAnalyzerBase.java:L66-L86):
    0 = MORE_THAN_ONE_MATCH (the reference then reports the read as not found), G_NO_SECOND = no second-best entry"""
    res = np.ascontiguousarray(res, dtype=GUIDED_RESULT)
    slices = np.ascontiguousarray(slices, dtype=np.uint8)
    anchor = np.ascontiguousarray(anchor, dtype=np.int32)
    n, stride = slices.shape
    assert res.shape == (n,) and anchor.shape == (n,)
    out = np.empty(n, dtype=np.int32)
    _check(gpu_lib().slr_guided_mismatch_diff(res.ctypes.data, n, slices.ctypes.data, stride, stride if slice_len is None else int(slice_len),
                                              anchor.ctypes.data, int(seq_len), C.byref(scores) if scores is not None else None, out.ctypes.data))
    return out


# ---------------------------------------------------------------------------------------------- UMI distances
class BestEditDistance:
    """Decoder of the packed int (ClusteringEditDistanceBase$BestEditDistance, java:L382-L416)."""
    MINUSONE, ZERO, PLUSONE = 0, 1, 2

    @staticmethod
    def getED(packed):
        return np.asarray(packed) & 0xFFFFFF

    @staticmethod
    def getPos1(packed):
        p = (np.asarray(packed).astype(np.int64) >> 27) & 7
        return np.where(p & 1, 0, np.where(p & 2, 1, 2))

    @staticmethod
    def getPos2(packed):
        p = (np.asarray(packed).astype(np.int64) >> 24) & 7
        return np.where(p & 1, 0, np.where(p & 2, 1, 2))


def out_offsets_for(job_offsets):
    sizes = np.diff(np.asarray(job_offsets, dtype=np.int64))
    oo = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes * sizes, out=oo[1:])
    return oo


def generate_distance_matrices(ctx, umis, job_offsets, umi_len=12, out=None, out_offsets=None):
    """ClusteringEditDistanceBase.generateDistanceMatrix for all jobs at once.
    umis u8 [m, stride] (umi_len+2 4-bit codes per read), job_offsets i64 [n_jobs+1] -> (flat int32, out_offsets)."""
    umis = np.ascontiguousarray(umis, dtype=np.uint8)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    if out_offsets is None:
        out_offsets = out_offsets_for(job_offsets)
    out_offsets = np.ascontiguousarray(out_offsets, dtype=np.int64)
    if out is None:
        out = np.empty(int(out_offsets[-1]), dtype=np.int32)
    _check(gpu_lib().slr_umi_dist(ctx.h, umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(job_offsets) - 1,
                                  out.ctypes.data, out_offsets.ctypes.data))
    return out, out_offsets


UMI_CLUSTER_REC = np.dtype([("n_neighbours", "<i4"), ("best_key", "<i4"), ("best_count", "<i4"), ("n_ties", "<i4")])


def cluster_local(ctx, umis, job_offsets, ed, umi_len=12, member=None, rank=None, want_matrices=False):
    """The two O(n^2) steps of ClusterOne_MyClustering.clusterLocal (ClusterOne_MyClustering.java:L175-L219) for all
    jobs at once, fused behind the distance matrices: per read |N(c)| and the chosen entry (first maximum of |N(l)|
    over the entries containing c; `rank` = the caller's iteration rank of the keys, None = ascending index, ties
    counted in n_ties).  Returns UMI_CLUSTER_REC [m] (+ the flat matrices and their offsets when asked)."""
    umis = np.ascontiguousarray(umis, dtype=np.uint8)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    m = int(job_offsets[-1]) if len(job_offsets) else 0
    rec = np.zeros(m, dtype=UMI_CLUSTER_REC)
    if member is not None:
        member = np.ascontiguousarray(member, dtype=np.uint8)
        if member.shape != (m,):
            raise ValueError("member must hold one byte per read")
    if rank is not None:
        rank = np.ascontiguousarray(rank, dtype=np.int32)
        if rank.shape != (m,):
            raise ValueError("rank must hold one int32 per read")
    out = out_offsets = None
    if want_matrices:
        out_offsets = out_offsets_for(job_offsets)
        out = np.empty(int(out_offsets[-1]), dtype=np.int32)
    _check(gpu_lib().slr_umi_cluster(ctx.h, umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(job_offsets) - 1,
                                     int(ed), member.ctypes.data if member is not None else None,
                                     rank.ctypes.data if rank is not None else None,
                                     out.ctypes.data if out is not None else None,
                                     out_offsets.ctypes.data if out_offsets is not None else None, rec.ctypes.data))
    return (rec, out, out_offsets) if want_matrices else rec


UMI_ASSIGN_REC = np.dtype([("center", "<i4"), ("u1", "i1"), ("u2", "i1"), ("pos2", "i1"), ("offset_center_mean", "i1"), ("flags", "<u2"),
                           ("cluster_size", "<u2"), ("n_clusters", "<i4")], align=True)
assert UMI_ASSIGN_REC.itemsize == 16
UA_ASSIGNED, UA_SKIPPED, UA_TIE_UNPIN, UA_DEEP = 1, 2, 4, 8


class UmiAssignParams(C.Structure):
    """The clustering knobs ClusterOneHierarchical reads (config.xml:270-278, UMIparameters.java:L96-L118, UmiClustering.java:L240)."""
    _fields_ = [("ed_complete", C.c_int32), ("ed_single", C.c_int32), ("single_threshold", C.c_int32), ("fold_depth", C.c_int32),
                ("max_hier", C.c_int32), ("deep", C.c_int32)]

    def __init__(self, umi_completelinkclusteringED=2, umi_singlelinkclusteringED=1, complexity_threshold_for_switch_to_single_link_clustering=3000,
                 foldDepthBelowMaxDiscardForClustering=50, max_hier=100, deep=1):
        super().__init__(umi_completelinkclusteringED, umi_singlelinkclusteringED, complexity_threshold_for_switch_to_single_link_clustering,
                         foldDepthBelowMaxDiscardForClustering, max_hier, deep)


def cluster_one_hierarchical(ctx, umis, job_offsets, umi_len=12, params=None, job_qv01=None, want_matrices=False):
    """ClusterOneHierarchical.call (ClusterOneHierarchical.java:L61-L217) for all jobs of at most 100 reads at once, fused behind the distance
    matrices: LingPipe complete link on the reads with a neighbour, the cut at umi_completelinkclusteringED, the depth rule, the cluster
    centres and, per read, the values setSamflagsAndStatsForClustered writes (centre, U1, U2, +-1 shift, mean shift of the cluster).
    Returns UMI_ASSIGN_REC [m] (+ the flat matrices and their offsets when asked)."""
    umis = np.ascontiguousarray(umis, dtype=np.uint8)
    job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
    m = int(job_offsets[-1]) if len(job_offsets) else 0
    rec = np.zeros(m, dtype=UMI_ASSIGN_REC)
    qv = None if job_qv01 is None else np.ascontiguousarray(job_qv01, dtype=np.uint8)
    if qv is not None and qv.shape != (len(job_offsets) - 1,):
        raise ValueError("job_qv01 must hold one byte per job")
    out = out_offsets = None
    if want_matrices:
        out_offsets = out_offsets_for(job_offsets)
        out = np.empty(int(out_offsets[-1]), dtype=np.int32)
    _check(gpu_lib().slr_umi_assign(ctx.h, umis.ctypes.data, umis.shape[1] if umis.ndim == 2 else 16, umi_len, job_offsets.ctypes.data,
                                    len(job_offsets) - 1, C.byref(params) if params is not None else None,
                                    qv.ctypes.data if qv is not None else None, out.ctypes.data if out is not None else None,
                                    out_offsets.ctypes.data if out_offsets is not None else None, rec.ctypes.data))
    return (rec, out, out_offsets) if want_matrices else rec


def split_oversized_group(n, ram_reserved):
    """UmiClustering.lambda$cluster$7 (UmiClustering.java:L136-L142): a (cell, region) group of n reads is cut into
    nChunks = ceil((float) n / sqrt(RAM_RESERVED / 300)) consecutive parts of n / nChunks + 1 reads (ListUtils.partition: the last part takes the
    rest) BEFORE it is clustered — the JVM's memory bound on one n x n matrix.  Caller-side: it changes the clusters, so a shadow class keeps it.
    Returns the part sizes."""
    import math
    max_square = int(ram_reserved) // 300
    n_chunks = int(math.ceil(float(np.float32(n)) / math.sqrt(float(max_square))))
    size = n // max(n_chunks, 1) + 1 if n_chunks > 0 else n + 1
    return [min(size, n - k) for k in range(0, n, size)]


def group_by_cell_and_region(cell_bc, region, valid=None, min_size=2):
    """UmiClustering.groupDataByCellAndRegion + the size filter of cluster() (UmiClustering.java:L97-L118, L135): the reads that have a cell barcode
    and a genomic region number, grouped by (barcode, region); groups of fewer than min_size reads are dropped.  Host-side (O(n log n), the caller's):
    returns (order, job_offsets) — order[job_offsets[j]:job_offsets[j + 1]] are the read indices of job j in INPUT order (the reference fills its
    lists from a parallel stream, so its order inside a group is arrival order; input order is the one-thread result), jobs in ascending
    (barcode, region) order (the reference iterates two ConcurrentHashMaps; the job order does not reach any per-read result)."""
    cell_bc = np.asarray(cell_bc, dtype=np.uint64)
    region = np.asarray(region, dtype=np.int64)
    idx = np.arange(len(cell_bc), dtype=np.int64) if valid is None else np.nonzero(np.asarray(valid, dtype=bool))[0].astype(np.int64)
    o = idx[np.lexsort((idx, region[idx], cell_bc[idx]))]
    if len(o) == 0:
        return o, np.zeros(1, dtype=np.int64)
    new = np.ones(len(o), dtype=bool)
    new[1:] = (cell_bc[o][1:] != cell_bc[o][:-1]) | (region[o][1:] != region[o][:-1])
    starts = np.nonzero(new)[0]
    sizes = np.diff(np.concatenate([starts, [len(o)]]))
    keep = np.repeat(sizes >= min_size, sizes)
    return o[keep], np.concatenate([[0], np.cumsum(sizes[sizes >= min_size])]).astype(np.int64)


class UmiSession:
    """The matrices of one batch of (cell, region) jobs kept on the device between calls (slr_umi_session_*): the distance kernels run
    once, cluster() can then be called with the caller's key order (`rank`) once the keys are known, and again on the unclustered
    subset (`member`) like ClusterOne_MyClustering.call does (…java:L73, L107)."""

    def __init__(self, ctx, umis, job_offsets, umi_len=12):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        self.job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
        self.n_reads = int(self.job_offsets[-1] - self.job_offsets[0]) if len(self.job_offsets) else 0
        h = C.c_void_p()
        _check(gpu_lib().slr_umi_session_create(ctx.h, umis.ctypes.data, umis.shape[1] if umis.ndim == 2 else 16, umi_len,
                                                self.job_offsets.ctypes.data, max(0, len(self.job_offsets) - 1), C.byref(h)))
        self.h = h

    def cluster(self, ed, member=None, rank=None):
        rec = np.zeros(self.n_reads, dtype=UMI_CLUSTER_REC)
        if member is not None:
            member = np.ascontiguousarray(member, dtype=np.uint8)
            assert member.shape == (self.n_reads,)
        if rank is not None:
            rank = np.ascontiguousarray(rank, dtype=np.int32)
            assert rank.shape == (self.n_reads,)
        _check(gpu_lib().slr_umi_session_cluster(self.h, int(ed), member.ctypes.data if member is not None else None,
                                                 rank.ctypes.data if rank is not None else None, rec.ctypes.data))
        return rec

    def assign(self, params=None, job_qv01=None):
        """ClusterOneHierarchical.call on the resident matrices (jobs of at most 100 reads): UMI_ASSIGN_REC per read"""
        rec = np.zeros(self.n_reads, dtype=UMI_ASSIGN_REC)
        qv = None if job_qv01 is None else np.ascontiguousarray(job_qv01, dtype=np.uint8)
        _check(gpu_lib().slr_umi_session_assign(self.h, C.byref(params) if params is not None else None,
                                                qv.ctypes.data if qv is not None else None, rec.ctypes.data))
        return rec

    def matrices(self):
        cells = int(gpu_lib().slr_umi_session_cells(self.h))
        out = np.empty(cells, dtype=np.int32)
        _check(gpu_lib().slr_umi_session_matrices(self.h, out.ctypes.data, cells))
        return out, out_offsets_for(self.job_offsets)

    def close(self):
        if self.h:
            gpu_lib().slr_umi_session_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def clusters_from_records(rec, job_offsets):
    """The grouping step the caller keeps (ClusterOne_MyClustering.java:L199, L219): per job the set of clusters,
    each the frozenset of keys that chose the same entry.  Jobs without any key give an empty set (Optional.empty)."""
    res = []
    for j in range(len(job_offsets) - 1):
        a, b = int(job_offsets[j]), int(job_offsets[j + 1])
        groups = {}
        for c in range(a, b):
            k = int(rec["best_key"][c])
            if k >= 0:
                groups.setdefault(k, set()).add(c - a)
        res.append({frozenset(v) for v in groups.values()})
    return res


# ---------------------------------------------------------------------------------------------- several GPUs, one caller
class MultiGpu:
    """All (or some) GPUs of the box behind one caller, as the single-JVM reference needs them (slr_multi_*): the barcode list is replicated,
    read batches and whole UMI jobs are dealt to the devices, results are positional, the BarcodesAssigned counters are summed on the device."""

    def __init__(self, n_devices=0, device_ids=None, n_streams=2):
        h = C.c_void_p()
        ids = None if device_ids is None else np.ascontiguousarray(device_ids, dtype=np.int32)
        _check(gpu_lib().slr_multi_create(int(n_devices if ids is None else len(ids)), None if ids is None else ids.ctypes.data, n_streams, C.byref(h)))
        self.h, self.table, self.keys = h, None, None

    @property
    def n_devices(self):
        return int(gpu_lib().slr_multi_n_devices(self.h))

    @property
    def peer_access(self):
        return bool(gpu_lib().slr_multi_peer_access(self.h))

    def load_barcodes(self, barcodes2bit, rank=None):
        self.keys = np.ascontiguousarray(barcodes2bit, dtype=np.uint64)
        r = None if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
        t = C.c_void_p()
        _check(gpu_lib().slr_multi_bc_table_create(self.h, self.keys.ctypes.data, None if r is None else r.ctypes.data, len(self.keys), 16, C.byref(t)))
        if self.table:
            gpu_lib().slr_multi_bc_table_destroy(self.table)
        self.table = t

    def assign_barcodes(self, slices, anchor, ed, plusminus=2, three_prime=True, lens=None, out=None):
        slices = np.ascontiguousarray(slices, dtype=np.uint8)
        anchor = np.ascontiguousarray(anchor, dtype=np.int32)
        n, stride = slices.shape
        if out is None:
            out = np.empty(n, dtype=BC_RESULT)
        lp = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.int32)
            lp = lens.ctypes.data
        _check(gpu_lib().slr_multi_bc_assign(self.h, self.table, int(ed), int(plusminus), int(three_prime), slices.ctypes.data, stride,
                                             min(stride, 32), lp, anchor.ctypes.data, n, out.ctypes.data))
        return out

    def counts(self):
        out = np.zeros((len(self.keys), 3), dtype=np.int64)
        _check(gpu_lib().slr_multi_bc_counts_read(self.h, self.table, out.ctypes.data))
        return out

    def reset_counts(self):
        _check(gpu_lib().slr_multi_bc_counts_reset(self.h, self.table))

    def umi_assign(self, umis, job_offsets, umi_len=12, params=None, job_qv01=None, out=None):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
        rec = np.zeros(int(job_offsets[-1]), dtype=UMI_ASSIGN_REC) if out is None else out
        qv = None if job_qv01 is None else np.ascontiguousarray(job_qv01, dtype=np.uint8)
        _check(gpu_lib().slr_multi_umi_assign(self.h, umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(job_offsets) - 1,
                                              C.byref(params) if params is not None else None, None if qv is None else qv.ctypes.data, rec.ctypes.data))
        return rec

    def umi_cluster(self, umis, job_offsets, ed, umi_len=12, member=None, rank=None):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
        rec = np.zeros(int(job_offsets[-1]), dtype=UMI_CLUSTER_REC)
        mb = None if member is None else np.ascontiguousarray(member, dtype=np.uint8)
        rk = None if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
        _check(gpu_lib().slr_multi_umi_cluster(self.h, umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(job_offsets) - 1, int(ed),
                                               None if mb is None else mb.ctypes.data, None if rk is None else rk.ctypes.data, rec.ctypes.data))
        return rec

    def umi_dist(self, umis, job_offsets, umi_len=12):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        job_offsets = np.ascontiguousarray(job_offsets, dtype=np.int64)
        oo = out_offsets_for(job_offsets)
        out = np.empty(int(oo[-1]), dtype=np.int32)
        _check(gpu_lib().slr_multi_umi_dist(self.h, umis.ctypes.data, umis.shape[1], umi_len, job_offsets.ctypes.data, len(job_offsets) - 1,
                                            out.ctypes.data, oo.ctypes.data))
        return out, oo

    def close(self):
        if getattr(self, "table", None):
            gpu_lib().slr_multi_bc_table_destroy(self.table)
            self.table = None
        if getattr(self, "h", None):
            gpu_lib().slr_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- cross-shard UMI merge
class UmiShardMerger:
    """Cross-shard merge of the UMI clustering jobs when the (cell, region)-sorted read stream is sharded by read index
    over the ranks: a (cell, region) group (UmiClustering.groupDataByCellAndRegion, UmiClustering.java:L97-L118) that
    straddles a shard boundary must be clustered as ONE job.  The rank that holds the group's first reads absorbs the
    leading reads of the following rank(s); those ranks drop them.  One all_gather of a few int64 per rank + one
    all_gather of the (padded) leading group of every rank — torch.distributed, NCCL on device tensors, gloo on the CPU.

    Keys are (cell barcode, region id) pairs; every rank passes the key of its first and last job."""

    META = 6          # first_cell, first_region, first_count, last_cell, last_region, n_jobs

    def __init__(self, cap=4096, group=None):
        self.cap, self.group = int(cap), group

    @staticmethod
    def plan(meta_all, rank, cap):
        """Pure part: meta_all [world, 6] (host ints).  Returns (give, takes) where give = number of leading reads this
        rank hands to a lower rank (its whole first job) and takes = [(source rank, count), ...] appended to its last job."""
        world = len(meta_all)
        fk = lambda r: (int(meta_all[r][0]), int(meta_all[r][1]))
        lk = lambda r: (int(meta_all[r][3]), int(meta_all[r][4]))
        nj = lambda r: int(meta_all[r][5])

        def gives(r):      # rank r's first job continues the previous non-empty rank's last job
            if r == 0 or nj(r) == 0:
                return False
            q = r - 1
            while q >= 0 and nj(q) == 0:
                q -= 1
            return q >= 0 and lk(q) == fk(r)
        give = int(meta_all[rank][2]) if gives(rank) else 0
        takes = []
        # a rank whose only job is handed away has nothing left to absorb into
        if nj(rank) > 0 and not (gives(rank) and nj(rank) == 1):
            r = rank + 1
            while r < world:
                if nj(r) == 0:
                    r += 1
                    continue
                if not (gives(r) and fk(r) == lk(rank)):
                    break
                takes.append((r, int(meta_all[r][2])))
                if nj(r) > 1:
                    break
                r += 1
        total = sum(c for _, c in takes)
        if total > cap or give > cap:
            raise SiceloreGpuError(SLR_E_UNSUPPORTED, "a (cell, region) group crossing a shard boundary has more than %d reads" % cap)
        return give, takes

    def exchange(self, umis, n_reads):
        """The data movement of the last planned merge alone (the plan is static while the sharding is): all_gather of every
        rank's leading group, the absorbed reads land behind this rank's own reads.  Returns the number of rows appended."""
        import torch
        import torch.distributed as dist
        first_count, give, takes = self.last_plan
        world = dist.get_world_size(self.group)
        stride = umis.shape[1]
        lead = torch.zeros((self.cap, stride), dtype=torch.uint8, device=umis.device)
        k = min(first_count, self.cap)
        if k:
            lead[:k] = umis[:k]
        lead_all = torch.empty(world * self.cap * stride, dtype=torch.uint8, device=umis.device)
        dist.all_gather_into_tensor(lead_all, lead.reshape(-1), group=self.group)
        lead_all = lead_all.reshape(world, self.cap, stride)
        extra = 0
        for src, cnt in takes:
            umis[n_reads + extra:n_reads + extra + cnt] = lead_all[src, :cnt]
            extra += cnt
        return extra

    def merge(self, umis, n_reads, job_offsets, first_key, last_key):
        """umis: torch uint8 [n_reads + cap, stride] (device or CPU; rows >= n_reads are free space), job_offsets: host int64
        [n_jobs + 1].  Returns (row0, n_rows, offsets): the jobs of this rank after the merge are `offsets` (host int64,
        relative to row0) over umis[row0 : row0 + n_rows]."""
        import torch
        import torch.distributed as dist
        job_offsets = np.asarray(job_offsets, dtype=np.int64)
        n_jobs = len(job_offsets) - 1
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        first_count = int(job_offsets[1] - job_offsets[0]) if n_jobs > 0 else 0
        meta = torch.tensor([first_key[0], first_key[1], first_count, last_key[0], last_key[1], n_jobs], dtype=torch.int64,
                            device=umis.device)
        meta_all = torch.empty(world * self.META, dtype=torch.int64, device=umis.device)
        dist.all_gather_into_tensor(meta_all, meta, group=self.group)
        meta_all = meta_all.cpu().numpy().reshape(world, self.META)
        give, takes = self.plan(meta_all, rank, self.cap)
        self.last_plan = (first_count, give, takes)
        extra = self.exchange(umis, n_reads)
        offs = job_offsets - job_offsets[0]
        row0 = 0
        if give:
            row0 = give
            offs = offs[1:] - give
        if extra:
            offs = offs.copy()
            offs[-1] += extra
        return row0, n_reads - row0 + extra, np.ascontiguousarray(offs, dtype=np.int64)
