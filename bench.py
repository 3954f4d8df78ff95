#!/usr/bin/env python
"""bench.py — throughput of the barcode + UMI edit-distance hot path on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle port, see below)
  python bench.py --workload umi5kx2k       # BASELINE.json configs[3]; bc737k_ed1 = configs[1]; default bc3m_ed2 = configs[2] (configs[4] at N = 8)

One step = one pass of the hot path over one batch per GPU (weak scaling: every rank owns its shard of the run's read stream):
  S1  slr_bc_assign   R reads x 5 window offsets vs the barcode list at --bcEditDistance                 (Parser.assignBarcode)
  S2  slr_umi_dist    the (cell, region) jobs of the shard -> packed 3x3 distance matrices, left in HBM  (generateDistanceMatrix)
  S6  slr_umi_assign  ClusterOneHierarchical on every job of <= 100 reads: clusters, centres, U1 / U2     (UMI assignment to SAM records)
  S5  slr_umi_cluster clusterLocal's two matrix passes on the deeper jobs (workload umi5kx2k only: the bc workloads have none)
  N > 1: the UMI read stream is the run's global (cell, region)-sorted stream cut by read index, so a job can straddle a shard boundary: it
  is clustered as ONE job on the lower rank (UmiShardMerger: all_gather of the boundary reads over NCCL, inside every step).  The
  per-barcode x ED counters (BarcodesAssigned.tsv) are all-reduced once per run, after the timed steps.
`value`  : reads/s with all inputs already resident in HBM, timed with CUDA events on the launching stream.
`e2e`    : the same step through the host-pointer C ABI (slr_bc_assign / slr_umi_assign) from pinned host buffers, H2D and D2H copies
           inside the timed region, a barrier before every step so that `ms_steps` means the same on every rank.
`roofline`: the dominant kernel against the INT-issue peak (warp instructions per unit from an ncu profile under profiles/ whose source
           hash equals the loaded library's, x units, / (SMs x 4 schedulers x SM clock x kernel time)); `roofline_hbm` = its measured DRAM
           bytes against MEASURED_PEAKS.json; `reference_equivalent_gbs` = SURVEY.md 8d's bytes of the REFERENCE algorithm (not a utilisation).
The reference arm cannot be the real thing: the path exists only as JVM bytecode and this image has no JVM.  It times the CPU oracle
(a restatement of the reference algorithm, oracle/, kind = "port") with every host thread on a bounded sample of the same workload.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/sec barcode+UMI assigned (ED<=2, 3M list) at 1/2/4/8 B200 vs host CPU"
WORKLOADS = {
    # name: (list size, list seed, read seed, bcEditDistance, kind)
    "bc3m_ed2": (3_000_000, 3_000_000, 2, 2, "bc"),       # BASELINE.json configs[2] — the configuration `metric` is quoted on; configs[4] at N = 8
    "bc737k_ed1": (737_280, 737, 1, 1, "bc"),             # configs[1]
    "umi5kx2k": (0, 0, 0, 0, "umi"),                      # configs[3]: UMI distance + clustering + assignment, 5 k cells x 2 k genes
}
UMI_SEED, UMI_MEAN, UMI_CAP = 4, 4.0, 2000
DEEP_JOB = 20_000                                          # the targeted-sequencing stress job of configs[3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bc3m_ed2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0 = 10 M; 12.5 M at 8 GPUs = the 100 M-read run of configs[4]; "
                                                         "40 M for umi5kx2k)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads of the CPU sample (0 = sized for ~20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-umi", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the NUMA node of its GPU")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML through pynvml (a query takes microseconds, so even a
    270 ms region of an 8-GPU run gets samples); `nvidia-smi` is the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]            # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.source = index, [], False, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv, self.source = None, "nvidia-smi"

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                    get = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    mask = int(get(self.h))
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if mask & b else "Not Active" for b in self.BITS])
                    time.sleep(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


class CudaArrayView:
    """zero-copy torch view of a raw device pointer (the table's counter buffer) via __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def bind_to_gpu_numa(local_rank):
    """Run this rank (and first-touch its pinned staging memory) on the NUMA node its GPU hangs off: at N = 8 every rank moves ~1 GB per step
    through the host, and a remote socket halves the link.  Returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as ex:
        return "not bound (%s)" % type(ex).__name__


def umi_jobs_for(pkg, n_reads, rank=0):
    """this rank's piece of the global (cell, region)-sorted read stream: reads [rank * n, (rank + 1) * n)"""
    return pkg.synth_umi_shard(rank * n_reads, n_reads, UMI_MEAN, UMI_CAP, UMI_SEED)


def umi5kx2k_batch(pkg, n_reads, rank):
    """configs[3]-shaped batch: geometric (cell, gene) jobs (mean 4, cap 2000) + one deep targeted-sequencing job of DEEP_JOB reads"""
    umis, offs, j0, j1 = pkg.synth_umi_shard(rank * n_reads, n_reads - DEEP_JOB, UMI_MEAN, UMI_CAP, UMI_SEED)
    du, doff = pkg.synth_umi_jobs(1, mean=1e9, cap=DEEP_JOB, seed=77 + rank)
    return np.concatenate([umis, du]), np.concatenate([offs, [offs[-1] + DEEP_JOB]]).astype(np.int64), j0, j1


def cpu_reference_step(orc, kind, bset, slices, anchor, ed, umis, offs, threads):
    t0 = time.perf_counter()
    res, probes = (None, 0)
    if kind == "bc":
        res, probes = orc.assign_barcode_batch(bset, slices, anchor, ed, 2, True, n_threads=threads)
    t1 = time.perf_counter()
    arec = None
    if umis is not None:
        mats, oo = orc.umi_matrix_batch(umis, offs, 12, n_threads=threads)
        arec = orc.umi_assign_batch(mats, offs, oo, None, None, n_threads=threads)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, probes, res, arec


def load_profile(pkg, kernel):
    """newest profiles/r*_kernel_profile.json entry for `kernel` measured on the sources this library was built from: the entry's own source hash
    (translation unit + transitive includes of that kernel, pkg.kernel_src_sha256) must match, or — older files — the hash over all of csrc/.
    None otherwise"""
    want_all, want_k = pkg.csrc_sha256(), pkg.kernel_src_sha256(kernel)
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernel_profile*.json")), reverse=True):
        try:
            prof = json.load(open(f))
        except Exception:
            continue
        for k in prof.get("kernels", []):
            if kernel != k.get("kernel", ""):
                continue
            if (k.get("src_sha256") and k["src_sha256"] == want_k) or (not k.get("src_sha256") and prof.get("csrc_sha256") == want_all):
                return dict(k, file=os.path.relpath(f, ROOT), csrc_sha256=k.get("src_sha256") or want_all)
    return None


def issue_roofline(pkg, kernel, units, ms, sm_mhz, sms, unit="read"):
    """INT-issue roofline of one kernel: warp instructions per unit (ncu count of the profile that matches the loaded sources) x units / kernel time,
    against SMs x 4 schedulers x SM clock.  Shared by bench.py and tools/bench_*.py so that every bench line of the repository names the bound
    that actually limits its kernel."""
    peak = sms * 4 * (sm_mhz or 1965.0) / 1e3
    roof = {"bound": "int_issue", "kernel": kernel, "kernel_ms_per_launch": ms, "units_per_launch": units, "unit": "Ginst/s", "peak": peak,
            "peak_source": "%d SMs x 4 schedulers x %.3f GHz (SM clock sampled under load)" % (sms, (sm_mhz or 1965.0) / 1e3)}
    prof = load_profile(pkg, kernel)
    if prof is not None and prof.get("inst_executed_per_unit"):
        ach = prof["inst_executed_per_unit"] * units / (ms / 1e3) / 1e9
        roof.update({"achieved": ach, "frac": ach / peak, "inst_per_" + unit: prof["inst_executed_per_unit"],
                     "traffic": (prof.get("dram_bytes_per_unit") or 0) * units or None, "profile": prof["file"], "profile_csrc_sha256": prof["csrc_sha256"]})
    else:
        roof.update({"achieved": None, "frac": None, "traffic": None,
                     "note": "no profile under profiles/ matches the loaded library's sources (csrc sha256 %s): run tools/make_profile.py" % pkg.csrc_sha256()[:16]})
    return roof, prof


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_wl, wl_seed, read_seed, ed, kind = WORKLOADS[a.workload]
    import __graft_entry__ as g
    pkg = g.load_package()
    threads = os.cpu_count() or 1
    R = a.reads or (40_000_000 if kind == "umi" else (12_500_000 if world == 8 else 10_000_000))
    use_umi = kind == "umi" or not a.no_umi
    if kind == "bc":
        wtxt = ("%s: %d synthetic 3' reads/GPU/step vs %d-barcode synthetic whitelist, bcEditDistance %d, testPlusMinusPos 2%s"
                % (a.workload, R, n_wl, ed, "" if not use_umi else " + UMI distance matrices, ClusterOneHierarchical clustering and per-read UMI "
                   "assignment of the same number of reads in (cell,region) jobs (geometric, mean 4)"))
    else:
        wtxt = ("umi5kx2k: %d reads/GPU/step in (cell,gene) jobs (geometric, mean 4, cap 2000) + one %d-read job: UMI distance matrices, "
                "ClusterOneHierarchical (jobs <= 100 reads) / ClusterOne_MyClustering (larger jobs), per-read UMI assignment" % (R, DEEP_JOB))
    config = {"workload": wtxt, "reads_per_gpu_per_step": R, "whitelist": n_wl, "bc_edit_distance": ed,
              "sharding": "reads sharded by index, list replicated; N>1: (cell,region) jobs cut by a shard boundary merged onto the lower rank "
                          "(all_gather over NCCL, every step); counters all-reduced once per run",
              "l2_policy": "inputs larger than L2 (%d MB of read slices / UMI codes per step) + table random access" % (R * 48 // 1_000_000)}

    # ------------------------------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        from oracle import orc
        bset = None
        if kind == "bc":
            wl = pkg.synth_whitelist(n_wl, wl_seed)
            bset = orc.BarcodeSet(wl, np.arange(1, n_wl + 1, dtype=np.int32))
            # bounded sample: calibrate on 2 000 reads, then size one step to ~4 s of host time
            cs, ca, _ = pkg.synth_reads(wl, 2000, seed=read_seed)
            cpu_reference_step(orc, kind, bset, cs[:200], ca[:200], ed, None, None, threads)
            tcal = cpu_reference_step(orc, kind, bset, cs, ca, ed, None, None, threads)[0]
            n_s = a.cpu_sample or int(min(2_000_000, max(2000, 4.0 * 2000 / tcal)))
            slices, anchor, _ = pkg.synth_reads(wl, n_s, seed=read_seed)
            umis, offs = (umi_jobs_for(pkg, n_s)[:2] if use_umi else (None, None))
        else:
            n_s = a.cpu_sample or 2_000_000
            slices = anchor = None
            umis, offs = umi5kx2k_batch(pkg, n_s, 0)[:2]
        ts = [cpu_reference_step(orc, kind, bset, slices, anchor, ed, umis, offs, threads)[0] for _ in range(a.steps)]
        t = sum(ts) / len(ts)
        v = n_s / t
        print(json.dumps({"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                          "data": "synthetic", "impl": "reference", "config": config,
                          "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port",
                                           "sample": "%d reads per step (bounded sample of the same workload), CPU oracle = restatement of the "
                                                     "reference's Java algorithm; the JVM path itself cannot run here" % n_s},
                          "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------------ our arm
    numa_note = "off" if a.no_numa else bind_to_gpu_numa(local_rank)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created (NCCL_DEBUG=VERSION on the box): fd 1 is pointed at
        # stderr until the JSON line is printed, so that stdout carries that one line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = pkg.Context(local_rank, n_streams=2)
    lib = pkg.gpu_lib()
    pin = lambda t: t.pin_memory()
    stream = torch.cuda.current_stream().cuda_stream
    table_build_ms = None
    if kind == "bc":
        wl = pkg.synth_whitelist(n_wl, wl_seed)
        t0 = time.perf_counter()
        table = pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, wl)
        table_build_ms = (time.perf_counter() - t0) * 1e3
        parser = pkg.Parser(ctx, table, bcEditDistance=ed, testPlusMinusPos=2, three_prime=True)
        # this rank's shard of the run: reads [rank*R, (rank+1)*R) — counter-based RNG, no communication
        h_slices = pin(torch.empty((R, 32), dtype=torch.uint8))
        h_anchor = pin(torch.empty(R, dtype=torch.int32))
        pkg.synth_reads(wl, R, seed=read_seed, first=rank * R, out=(h_slices.numpy(), h_anchor.numpy()))
        h_res = pin(torch.empty((R, 32), dtype=torch.uint8))
        d_slices, d_anchor = h_slices.to(dev), h_anchor.to(dev)
        d_res = torch.empty((R, 32), dtype=torch.uint8, device=dev)
        cptr, cn = table.counts_device_ptr()
        d_counts = torch.as_tensor(CudaArrayView(cptr, cn), device=dev)
        np_res = h_res.numpy().view(pkg.BC_RESULT).reshape(-1)
    MERGE_CAP = 4096
    merged = None
    if use_umi:
        umis_np, offs_np, job_first, job_last = (umi5kx2k_batch(pkg, R, rank) if kind == "umi" else umi_jobs_for(pkg, R, rank))
        n_own = len(umis_np)
        h_umis_all = pin(torch.zeros((n_own + MERGE_CAP, 16), dtype=torch.uint8))
        h_umis_all[:n_own] = torch.from_numpy(umis_np)
        d_umis_all = h_umis_all.to(dev)
        row0, n_rows, moffs = 0, n_own, offs_np
        merger = None
        if world > 1 and kind == "bc":
            # cross-shard UMI merge: the job cut by a shard boundary is clustered as ONE job on the lower rank; the keys are the jobs' ids in
            # the global stream (cell = id // 2000, region = id % 2000)
            merger = pkg.UmiShardMerger(cap=MERGE_CAP)
            row0, n_rows, moffs = merger.merge(d_umis_all, n_own, offs_np, (job_first // 2000, job_first % 2000), (job_last // 2000, job_last % 2000))
            merged = (merger.last_plan[1], list(merger.last_plan[2]))
        moo = pkg.out_offsets_for(moffs)
        n_jobs, dev_cells = len(moffs) - 1, int(moo[-1])
        h_moffs = pin(torch.from_numpy(np.ascontiguousarray(moffs)))
        d_umis, d_offs, d_oo = d_umis_all[row0:], h_moffs.to(dev), torch.from_numpy(moo).to(dev)
        d_mat = torch.empty(dev_cells, dtype=torch.int32, device=dev)
        d_arec = torch.empty((n_rows, 16), dtype=torch.uint8, device=dev)
        # job lists + the working arrays of the jobs above 100 reads (ClusterOne_MyClustering's, clustered in the same call)
        n_deep_jobs = int((np.diff(moffs) > 100).sum())
        ascr_bytes = int(lib.slr_umi_assign_scratch_bytes(n_jobs)) + sum(int(lib.slr_umi_assign_deep_job_bytes(int(n))) for n in np.diff(moffs) if n > 100)
        d_ascr = torch.empty(ascr_bytes, dtype=torch.uint8, device=dev)
        h_arec = pin(torch.empty((n_rows, 16), dtype=torch.uint8))
    n_units = R

    def step_device(ev=None):
        if ev:
            ev[0].record()
        if kind == "bc":
            parser.assign_barcodes_dev(d_slices.data_ptr(), 32, d_anchor.data_ptr(), R, d_res.data_ptr(), stream)
        if ev:
            ev[1].record()
        if use_umi:
            if merger is not None:
                merger.exchange(d_umis_all, n_own)        # NCCL all_gather of the boundary jobs' reads
            pkg._check(lib.slr_umi_dist_dev(ctx.h, d_umis.data_ptr(), 16, 12, d_offs.data_ptr(), n_jobs, n_rows, d_mat.data_ptr(),
                                            d_oo.data_ptr(), dev_cells, stream))
            if ev:
                ev[2].record()
            if ev:
                ev[3].record()
            pkg._check(lib.slr_umi_assign_dev2(ctx.h, d_mat.data_ptr(), d_offs.data_ptr(), d_oo.data_ptr(), n_jobs, n_rows, None, None,
                                               d_ascr.data_ptr(), ascr_bytes, d_arec.data_ptr(), stream))
        if ev:
            ev[4].record()

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=2)
    h_umis_m = h_umis_all[row0:row0 + n_rows] if use_umi else None

    def e2e_umi():
        if merger is not None:                            # the absorbed boundary reads reach the host buffer behind the rank's own reads
            extra = merger.exchange(d_umis_all, n_own)
            if extra:
                h_umis_all[n_own:n_own + extra].copy_(d_umis_all[n_own:n_own + extra])
        arec = h_arec.numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)
        pkg._check(lib.slr_umi_assign(ctx.h, h_umis_m.data_ptr(), 16, 12, h_moffs.data_ptr(), n_jobs, None, None, None, None, arec.ctypes.data))

    def step_e2e():
        # the two seams are independent calls of two host threads (the reference's worker pools are concurrent too); each call copies its
        # inputs H2D, runs its kernels and copies its records D2H before it returns; the matrices stay on the device (records only)
        f = pool.submit(parser.assign_barcodes, h_slices.numpy(), h_anchor.numpy(), None, np_res) if kind == "bc" else None
        if use_umi:
            e2e_umi()
        if f is not None:
            f.result()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_device()
    sync_all()
    if kind == "bc":
        table.reset_counts()                                  # the counters then hold exactly the K timed steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = pkg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    legs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(a.steps)]
    e0.record()
    for s in range(a.steps):
        step_device(legs[s])
    e1.record()
    sync_all()
    launches = pkg.launch_count() - l0
    clocks = sampler.summary()                             # stop sampling here: nvidia-smi takes driver locks that stall the
    sampler.join(timeout=10)                               # host-side submission of the end-to-end leg below
    ms_total = e0.elapsed_time(e1)
    bc_ms = sum(x[0].elapsed_time(x[1]) for x in legs) / a.steps
    dist_ms = sum(x[1].elapsed_time(x[2]) for x in legs) / a.steps if use_umi else 0.0
    cluster_ms = sum(x[2].elapsed_time(x[3]) for x in legs) / a.steps if use_umi else 0.0
    assign_ms = sum(x[3].elapsed_time(x[4]) for x in legs) / a.steps if use_umi else 0.0

    # ---- parity on THIS rank (every rank): a strided sample of the barcode records and of the UMI jobs against the oracle; the job absorbed
    # ---- across a shard boundary against the oracle on the whole job as the unsharded run sees it; the all-reduced counters
    parity = {"bc_sample": None, "umi_sample": None, "boundary_job": None, "counters": None}
    if not a.no_cpu_baseline:
        from oracle import orc
        if kind == "bc":
            sel = np.arange(0, R, max(1, R // (4000 if ed >= 2 else 40000)))
            bset = orc.BarcodeSet(wl, np.arange(1, n_wl + 1, dtype=np.int32))
            exp, _ = orc.assign_barcode_batch(bset, h_slices.numpy()[sel], h_anchor.numpy()[sel], ed, 2, True, n_threads=threads)
            got = d_res.cpu().numpy().view(pkg.BC_RESULT).reshape(-1)
            parity["bc_sample"] = bool((got[sel] == exp).all())
            parity["bc_sample_reads"] = int(len(sel))
        if use_umi:
            h_m = d_umis.cpu().numpy()[:n_rows]
            arec = d_arec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)
            jsel = np.unique(np.concatenate([np.arange(0, n_jobs, max(1, n_jobs // 20000)), [0, n_jobs - 1]]))
            jsel = np.unique(np.concatenate([jsel, np.nonzero(np.diff(moffs) > 100)[0][:64]]))     # + (up to 64 of) the jobs ClusterOne_MyClustering gets
            ok, reads_checked = True, 0
            sub_u = np.concatenate([h_m[moffs[j]:moffs[j + 1]] for j in jsel])
            sub_o = np.concatenate([[0], np.cumsum([moffs[j + 1] - moffs[j] for j in jsel])]).astype(np.int64)
            em, eoo = orc.umi_matrix_batch(sub_u, sub_o, 12, n_threads=threads)
            erec = orc.umi_assign_batch(em, sub_o, eoo, None, None, n_threads=threads)
            gmat = d_mat.cpu().numpy() if dev_cells < 4_000_000_000 else None
            for k, j in enumerate(jsel):
                n_j = int(moffs[j + 1] - moffs[j])
                ok &= arec[moffs[j]:moffs[j + 1]].tobytes() == erec[sub_o[k]:sub_o[k + 1]].tobytes()
                if gmat is not None:
                    ok &= bool((gmat[moo[j]:moo[j] + n_j * n_j] == em[eoo[k]:eoo[k + 1]]).all())
                reads_checked += n_j
            parity["umi_sample"] = bool(ok)
            parity["umi_sample_reads"] = int(reads_checked)
            if merged is not None and merged[1]:
                # this rank absorbed the head of the next shard(s): its last job must equal the global job `job_last` as a whole
                gu, go = pkg.synth_umi_jobs_at(job_last, 1, UMI_MEAN, UMI_CAP, UMI_SEED)
                j = n_jobs - 1
                have = h_m[moffs[j]:moffs[j + 1]]
                tail_ok = len(gu) >= len(have) and bool((gu[len(gu) - len(have):] == have).all())   # (its own head may lie on the rank before)
                fm, foo = orc.umi_matrix_batch(have, np.array([0, len(have)], dtype=np.int64), 12)
                frec = orc.umi_assign_batch(fm, np.array([0, len(have)], dtype=np.int64), foo)
                parity["boundary_job"] = bool(tail_ok and arec[moffs[j]:moffs[j + 1]].tobytes() == frec.tobytes() and
                                              (gmat is None or (gmat[moo[j]:moo[j + 1]] == fm).all()))
                parity["boundary_job_reads"] = int(len(have))
    if kind == "bc":
        # BarcodesAssigned.tsv counters: all-reduced ONCE per run; = steps x the histogram of this run's assigned reads, summed over the ranks
        got = d_res.cpu().numpy().view(pkg.BC_RESULT).reshape(-1)
        okm = (got["flags"] & 1) != 0
        hist = torch.zeros(3, dtype=torch.int64, device=dev)
        hist += torch.from_numpy(np.bincount(got["ed"][okm], minlength=3)[:3].astype(np.int64)).to(dev)
        if world > 1:
            dist.all_reduce(d_counts)                     # cross-shard merge of the counters (sum, int64, NCCL)
            dist.all_reduce(hist)
        torch.cuda.synchronize()
        csum = d_counts.view(-1, 3).sum(dim=0)
        parity["counters"] = bool((csum == hist * a.steps).all().item())
        assigned = int(okm.sum())
    pvals = [v for k, v in parity.items() if k in ("bc_sample", "umi_sample", "boundary_job", "counters") and v is not None]
    parity_rank = bool(all(pvals)) if pvals else None

    # host <-> device link of this box (pinned, 256 MB each way): explains e2e on a slow PCIe slot / remote NUMA node
    link = {}
    pb, db = pin(torch.empty(256 << 20, dtype=torch.uint8)), torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, dst, src in (("h2d_gbs", db, pb), ("d2h_gbs", pb, db)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link[name] = (256 << 20) / (c0.elapsed_time(c1) / 1e3) / 1e9
    del pb, db
    # end to end through the host-pointer ABI
    # warm-up: the context hands its stream slots out round-robin and every slot grows its own staging buffers on first
    # use, so each seam is called once per slot on its own before the W concurrent warm-up steps
    for _ in range(2 if kind == "bc" else 0):
        parser.assign_barcodes(h_slices.numpy(), h_anchor.numpy(), None, np_res)
    for _ in range(2 if use_umi else 0):
        e2e_umi()
    for _ in range(a.warmup):
        step_e2e()
    sync_all()
    e2e_steps = []
    for _ in range(a.steps):
        if world > 1:
            dist.barrier()                                # every rank starts the step together: ms_steps means the same on every rank
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        step_e2e()
        e2e_steps.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    # the host-pointer calls must have produced the very records the device-resident step left in HBM (same inputs, same library)
    e2e_same = True
    if kind == "bc":
        e2e_same &= bool((np_res == d_res.cpu().numpy().view(pkg.BC_RESULT).reshape(-1)).all())
    if use_umi:
        e2e_same &= bool((h_arec.numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)[:n_rows] == d_arec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)).all())
    parity["e2e_matches_device"] = e2e_same
    if parity_rank is not None:
        parity_rank = parity_rank and e2e_same
    # per step the slowest rank counts; the step time of the run is the mean of those maxima
    t_steps = torch.tensor(e2e_steps, dtype=torch.float64, device=dev)
    t = torch.tensor([ms_total / a.steps, bc_ms, dist_ms, assign_ms, 0.0 if parity_rank is False else 1.0, cluster_ms], dtype=torch.float64, device=dev)
    lk = torch.tensor([link["h2d_gbs"], link["d2h_gbs"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_steps, op=dist.ReduceOp.MAX)
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        t[4] = tmin[4]
        dist.all_reduce(lk, op=dist.ReduceOp.MIN)
    e2e_steps = [round(float(x), 2) for x in t_steps.cpu()]
    e2e_ms = float(t_steps.mean())
    ms_step, bc_ms, dist_ms, assign_ms, parity_all, cluster_ms = (float(x) for x in t.cpu())

    if rank == 0:
        out = {"metric": METRIC, "value": world * n_units / (ms_step / 1e3), "unit": "reads/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
               "legs_ms": {"bc_assign": bc_ms, "umi_dist": dist_ms, "umi_assign": assign_ms + cluster_ms},
               "parity_all_ranks": (None if a.no_cpu_baseline else bool(parity_all >= 1.0)), "parity_rank0": parity, "numa": numa_note}
        if kind == "bc":
            out["assigned_fraction"] = assigned / R
            out["table_build_ms"] = table_build_ms
        h2d = (R * 36 if kind == "bc" else 0) + (n_rows * 16 + 8 * (n_jobs + 1) if use_umi else 0)
        d2h = (R * 32 if kind == "bc" else 0) + (n_rows * 16 if use_umi else 0)
        out["e2e"] = {"value": world * n_units / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": e2e_ms, "ms_steps": e2e_steps, "link_min_over_ranks": {"h2d_gbs": float(lk[0]), "d2h_gbs": float(lk[1])},
                      "results": "records only (32 B/read barcode records, 16 B/read UMI assignment records); matrices stay in HBM"}
        # ---- CPU baseline (bounded sample, all host threads) + the reference's algorithmic bytes per read -------------
        probes_per_read = 55091.0 if ed >= 2 else 620.0       # App. A.5 of SURVEY.md; re-measured on the sample below
        if not a.no_cpu_baseline:
            from oracle import orc
            if kind == "bc":
                n_s = min(a.cpu_sample or (30_000 if ed >= 2 else 2_000_000), R)
                su, so = (umi_jobs_for(pkg, n_s)[:2] if use_umi else (None, None))
                sl, an = h_slices.numpy()[:n_s], h_anchor.numpy()[:n_s]
                tt, tbc, probes, cres, _ = cpu_reference_step(orc, kind, bset, sl, an, ed, su, so, threads)
                probes_per_read = probes / n_s
                gres = d_res[:n_s].cpu().numpy().view(pkg.BC_RESULT).reshape(-1)
                out["cpu_baseline"] = {"value": n_s / tt, "unit": "reads/s", "cores": threads, "kind": "port",
                                       "sample": "first %d reads of the step's batch (+ the UMI jobs of as many reads); CPU oracle = restatement "
                                                 "of the reference's Java algorithm, OpenMP over reads / jobs" % n_s,
                                       "bc_only_reads_per_s": n_s / tbc, "probes_per_read": probes_per_read,
                                       "gpu_matches_oracle_on_sample": bool((gres == cres).all())}
            else:
                n_s = min(a.cpu_sample or 2_000_000, R)
                su, so = umi5kx2k_batch(pkg, n_s, 0)[:2]
                tt = cpu_reference_step(orc, kind, None, None, None, 0, su, so, threads)[0]
                out["cpu_baseline"] = {"value": n_s / tt, "unit": "reads/s", "cores": threads, "kind": "port",
                                       "sample": "a %d-read batch of the same shape (incl. the %d-read job); CPU oracle, OpenMP over jobs" % (n_s, DEEP_JOB)}
        # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_ghz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) / 1e3
        issue_peak = sms * 4 * sm_ghz                          # G warp instructions / s: 4 schedulers per SM, one issue per cycle each
        # the dominant KERNEL: the assign leg of the umi workload is two kernels (umi_assign_kernel for the small jobs, umi_assign_deep_kernel for
        # the 20 000-read job, whose time is single-thread container emulation: profiles/r2_deep_timings.txt), so it does not compete as one
        if kind == "bc":
            cands = [("bc_assign_kernel<%d>" % ed, bc_ms, n_units, "read"), ("umi_pairs_kernel", dist_ms, n_units, "read"),
                     ("umi_assign_kernel", assign_ms, n_units, "read")]
        else:                                              # per read PAIR: the batch holds one job with 2 x 10^8 of them
            cands = [("umi_pairs_kernel/pair", dist_ms, int(((np.diff(moffs) * (np.diff(moffs) - 1)) // 2).sum()), "pair")]
        dom = max(cands, key=lambda x: x[1])
        kname, kms, kunits, unit = dom
        prof = load_profile(pkg, kname)
        roof = {"bound": "int_issue", "kernel": kname, "kernel_ms_per_launch": kms, "units_per_launch": kunits, "unit": "Ginst/s", "peak": issue_peak,
                "peak_source": "%d SMs x 4 schedulers x %.3f GHz (SM clock sampled under load)" % (sms, sm_ghz)}
        if prof is not None and prof.get("inst_executed_per_unit"):
            ach = prof["inst_executed_per_unit"] * kunits / (kms / 1e3) / 1e9
            roof.update({"achieved": ach, "frac": ach / issue_peak, "inst_per_" + unit: prof["inst_executed_per_unit"],
                         "traffic": (prof.get("dram_bytes_per_unit") or 0) * kunits or None, "profile": prof["file"],
                         "profile_csrc_sha256": prof["csrc_sha256"], "lib_sha256": pkg.lib_sha256()})
            if prof.get("dram_bytes_per_unit"):
                gbs = prof["dram_bytes_per_unit"] * kunits / (kms / 1e3) / 1e9
                out["roofline_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                       "traffic": prof["dram_bytes_per_unit"] * kunits, "compulsory_bytes_per_" + unit: 68 if kname.startswith("bc") else (4 if unit == "pair" else 36),
                                       "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}
        else:
            roof.update({"achieved": None, "frac": None, "traffic": None,
                         "note": "no profile under profiles/ matches the loaded library's sources (csrc sha256 %s): run tools/make_profile.sh" % pkg.csrc_sha256()[:16]})
        out["roofline"] = roof
        if kind == "bc":
            bytes_per_read = probes_per_read * 8 + 36 + 32      # reference probes x 8 B key + boundary in / out
            out["reference_equivalent_gbs"] = {"value": bytes_per_read * R / (bc_ms / 1e3) / 1e9, "algorithmic_bytes_per_read": bytes_per_read,
                                               "note": "bytes the REFERENCE algorithm would touch for the same reads (hash probes x 8 B + boundary "
                                                       "in/out), divided by the kernel time: a work-equivalence figure, not a utilisation"}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
