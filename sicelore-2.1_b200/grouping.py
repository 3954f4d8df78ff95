"""Host side of the clustering seam (first half): forming the (cell, region) jobs `slr_umi_assign` is fed with, for callers
without a JVM.  Thin ctypes mirror of the reference's classes over the library's native host code (csrc/slr_group.cpp):

* `ReadGrouper.group_sams`  — `ReadGrouper.groupSams` (F!com/rw/umifinder/bamreaders/ReadGrouper.class, ReadGrouper.java:L82-L230) -> `slr_grouper_group_sams`
* `group_stream`            — the chunk loop of `BamReader.run` (BamReader.java:L106-L145) around it
* `group_jobs`              — `UmiClustering.cluster` up to its Submitter (UmiClustering.java:L97-L118, L134-L143) -> `slr_group_jobs`

No device is involved (the reference does this work on its reader thread: O(n log n) over positions and keys, no edit distances); the
functions work without a GPU.  Parity: tests/golden/ref_grouper.npz (400 groupSams calls) and ref_jobs.npz (80 cluster() calls), both produced
by the reference's own class files (oracle/make_ref_grouper.py, oracle/make_ref_jobs.py)."""
import ctypes as C

import numpy as np

from . import SLR_E_REFERENCE_THROWS, SiceloreGpuError, _check, gpu_lib


class NullCenterError(SiceloreGpuError):
    """SLR_E_REFERENCE_THROWS from slr_grouper_group_sams: java.lang.NullPointerException at ReadGrouper.java:L173"""


class ReadGrouper:
    """One instance = the process-wide state of the reference class: `MAX_GENOME_DISTANCE_FOR_SAME_GENOMIC_REGION`
    (config.xml:247 max_GenomeDistance_forGrouping, default 500) and the static region counter (ReadGrouper.java:L460)."""

    def __init__(self, max_genome_distance=500, first_region_id=0):
        self.h = C.c_void_p()
        _check(gpu_lib().slr_grouper_create(int(max_genome_distance), int(first_region_id), C.byref(self.h)))

    def close(self):
        if self.h:
            gpu_lib().slr_grouper_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    @property
    def next_region_id(self):
        return int(gpu_lib().slr_grouper_next_region_id(self.h))

    def group_sams(self, position, flags, region, keep_data_end, has_position=None):
        """One chunk of SAM records in BAM order (arguments as slr_grouper_group_sams, include/sicelore_host.h).  `region` (int64, in / out):
        -1 = no region number.  Returns last_index (None for an empty chunk): reads [0, last_index] are the grouped chunk, the reads behind
        it open the next chunk when keep_data_end."""
        p64 = np.asarray(position, dtype=np.int64)
        if len(p64) and (p64.min() < -(1 << 31) or p64.max() >= (1 << 31)):
            raise SiceloreGpuError(-1, "group_sams: positions are Java ints (SAM coordinates), got a value outside int32")
        position = np.ascontiguousarray(p64, dtype=np.int32)
        flags = np.ascontiguousarray(np.asarray(flags, dtype=np.int64) & 0xFFFF, dtype=np.int32)
        n = len(position)
        assert flags.shape == (n,) and region.shape == (n,) and region.dtype == np.int64 and region.flags.c_contiguous
        has = None if has_position is None else np.ascontiguousarray(has_position, dtype=np.uint8)
        assert has is None or has.shape == (n,)
        last = C.c_int64(-1)
        rc = gpu_lib().slr_grouper_group_sams(self.h, position.ctypes.data, None if has is None else has.ctypes.data, flags.ctypes.data, n,
                                              int(bool(keep_data_end)), region.ctypes.data, C.byref(last))
        if rc == SLR_E_REFERENCE_THROWS:
            raise NullCenterError(rc, gpu_lib().slr_last_error().decode("utf-8", "replace"))
        _check(rc)
        return None if n == 0 else int(last.value)


def group_stream(grouper, position, flags, chromosome, chunk_size, has_position=None):
    """BamReader.run's loop (BamReader.java:L106-L145) over a whole coordinate-sorted record stream: a chunk is grouped when chunk_size records
    have been read since the last grouping (keepDataEnd = true: its tail is carried into the next chunk) or when the reference name changes
    (keepDataEnd = false), the record that triggered the grouping opens the next chunk; what is left at the end is grouped with
    keepDataEnd = false.  chromosome[i] = any value comparable with == (reference index or name).  Returns (region, emitted): region[i] = the
    region number record i carried when its chunk was handed on (-1 = none), emitted = the list of index arrays, one per grouped chunk, in
    the order the clustering stage receives them.  (The loop is restated from the bytecode; it needs htsjdk and is not part of the
    interpreter-pinned vectors — `group_sams` is.)"""
    position = np.asarray(position, dtype=np.int64)
    flags = np.asarray(flags, dtype=np.int64)
    n = len(position)
    has = np.ones(n, dtype=bool) if has_position is None else np.asarray(has_position, dtype=bool)
    region = np.full(n, -1, dtype=np.int64)
    emitted = []
    if n == 0:
        return region, emitted

    def group(ids, keep):
        ids = np.asarray(ids, dtype=np.int64)
        reg = np.ascontiguousarray(region[ids])
        li = grouper.group_sams(position[ids], flags[ids], reg, keep, has[ids])
        region[ids] = reg
        emitted.append(ids[:li + 1])
        return list(ids[li + 1:]) if keep else []

    chunk = [0]                                                # L116-L121
    chrom = chromosome[0]
    counter = 1
    for i in range(1, n):                                      # L126-L143
        counter += 1
        end_of_chromosome = chromosome[i] != chrom
        if end_of_chromosome:
            chrom = chromosome[i]
        if counter >= chunk_size or end_of_chromosome:
            chunk = group(chunk, not end_of_chromosome)
            counter = 0
        chunk.append(i)
    if chunk:                                                  # L144-L145
        group(chunk, False)
    return region, emitted


def group_jobs(cell_bc, region, valid=None, min_size=2, ram_reserved=0):
    """UmiClustering.cluster up to its Submitter (slr_group_jobs): (order, job_offsets) — order[job_offsets[j]:job_offsets[j + 1]] are the read
    indices of job j in input order; ram_reserved != 0 applies the reference's split of oversized groups (RAM_RESERVED, UmiClustering.java:L59)."""
    cell_bc = np.ascontiguousarray(cell_bc, dtype=np.uint64)
    region = np.ascontiguousarray(region, dtype=np.int64)
    n = len(cell_bc)
    assert region.shape == (n,)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    assert v is None or v.shape == (n,)
    order = np.zeros(max(n, 1), dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    nj = C.c_int64(0)
    _check(gpu_lib().slr_group_jobs(cell_bc.ctypes.data, region.ctypes.data, None if v is None else v.ctypes.data, n, int(min_size),
                                    int(ram_reserved), order.ctypes.data, off.ctypes.data, C.byref(nj)))
    k = int(nj.value)
    return order[:int(off[k])].copy(), off[:k + 1].copy()
