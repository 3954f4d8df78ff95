"""Measurement of the UMI rows (SURVEY.md §8 a12-a14 + §8f-3) on BASELINE.json's configs[3] shape — "UMI assignment + clustering, 5k cells x
2k genes": (cell, gene) jobs of geometric size (mean 4, cap 2 000) plus one targeted-sequencing job — same shape as bench.py's line:
device-resident throughput (CUDA events over slr_umi_dist_dev + slr_umi_cluster_dev), the host-pointer C-ABI call with H2D / D2H inside the
timed region (`e2e`: slr_umi_cluster returning matrices and records; `e2e_records_only`: matrices left on the device), the CPU oracle on a
bounded sample (`cpu_baseline`) and the byte bookkeeping.
One B200:   python tools/bench_umi_cluster.py [--jobs 10000000] [--deep 20000] [--ed 2] [--steps 5] [--warmup 3]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
from bench import ClockSampler

ap = argparse.ArgumentParser()
ap.add_argument("--jobs", type=int, default=10_000_000)
ap.add_argument("--deep", type=int, default=20_000)
ap.add_argument("--ed", type=int, default=2)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu-jobs", type=int, default=400_000)
a = ap.parse_args()
pkg = g.load_package()
from oracle import orc
orc.build()

umis, offs = pkg.synth_umi_jobs(a.jobs, mean=4.0, cap=2000, seed=4)
if a.deep > 0:                                                   # the targeted-sequencing stress job goes last
    du, _ = pkg.synth_umi_jobs(1, mean=1e9, cap=a.deep, seed=8)
    umis = np.concatenate([umis, du])
    offs = np.concatenate([offs, [offs[-1] + len(du)]]).astype(np.int64)
oo = pkg.out_offsets_for(offs)
m, cells, n_jobs = len(umis), int(oo[-1]), len(offs) - 1
ctx = pkg.Context(0, 2)
dev = torch.device("cuda", 0)
pin = lambda x: torch.from_numpy(x).pin_memory()
h_u, h_o, h_oo = pin(umis), pin(offs), pin(oo)
d_u, d_o, d_oo = (t.to(dev) for t in (h_u, h_o, h_oo))
d_m = torch.empty(cells, dtype=torch.int32, device=dev)
d_cnt = torch.empty(m, dtype=torch.int32, device=dev)
d_rec = torch.empty(m * 4, dtype=torch.int32, device=dev)
h_m = torch.empty(cells, dtype=torch.int32).pin_memory()
h_rec = torch.empty(m * 4, dtype=torch.int32).pin_memory()
lib, st = pkg.gpu_lib(), torch.cuda.current_stream().cuda_stream


def step_device():
    pkg._check(lib.slr_umi_dist_dev(ctx.h, d_u.data_ptr(), 16, 12, d_o.data_ptr(), n_jobs, m, d_m.data_ptr(), d_oo.data_ptr(), cells, st))
    pkg._check(lib.slr_umi_cluster_dev(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), n_jobs, m, a.ed, None, None, d_cnt.data_ptr(),
                                       d_rec.data_ptr(), st))


def step_e2e(with_matrices):
    pkg._check(lib.slr_umi_cluster(ctx.h, h_u.data_ptr(), 16, 12, h_o.data_ptr(), n_jobs, a.ed, None, None,
                                   h_m.data_ptr() if with_matrices else None, h_oo.data_ptr() if with_matrices else None, h_rec.data_ptr()))


def timed_events(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for _ in range(a.warmup):
    step_device()
torch.cuda.synchronize()
l0 = pkg.launch_count()
sampler = ClockSampler(0)
sampler.start()
ms = timed_events(step_device, a.steps)
clocks = sampler.summary()
sampler.join(timeout=10)
launches = pkg.launch_count() - l0
ms_dist = timed_events(lambda: pkg._check(lib.slr_umi_dist_dev(ctx.h, d_u.data_ptr(), 16, 12, d_o.data_ptr(), n_jobs, m, d_m.data_ptr(),
                                                                d_oo.data_ptr(), cells, st)), a.steps)
rec_dev = d_rec.cpu().numpy().view(pkg.UMI_CLUSTER_REC).reshape(m)
e2e = {}
for key, wm in (("e2e", True), ("e2e_records_only", False)):
    for _ in range(max(2, a.warmup)):
        step_e2e(wm)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e(wm)
    t = (time.perf_counter() - t0) / a.steps
    e2e[key] = {"value": m / t, "unit": "reads/s", "ms_per_step": t * 1e3, "h2d_bytes_per_step": m * 16 + (n_jobs + 1) * 8,
                "d2h_bytes_per_step": m * 16 + (cells * 4 if wm else 0)}
# the two-call protocol of a session: distance kernels once, cluster without and then with the caller's key order (a stand-in order here)
h_rank = pin((np.arange(m, dtype=np.int64) - np.repeat(offs[:-1], np.diff(offs))).astype(np.int32)[::1].copy())
h_rec2 = torch.empty(m * 4, dtype=torch.int32).pin_memory()
import ctypes as C


def step_session():
    h = C.c_void_p()
    pkg._check(lib.slr_umi_session_create(ctx.h, h_u.data_ptr(), 16, 12, h_o.data_ptr(), n_jobs, C.byref(h)))
    pkg._check(lib.slr_umi_session_cluster(h, a.ed, None, None, h_rec2.data_ptr()))
    pkg._check(lib.slr_umi_session_cluster(h, a.ed, None, h_rank.data_ptr(), h_rec2.data_ptr()))
    lib.slr_umi_session_destroy(h)


for _ in range(2):
    step_session()
t0 = time.perf_counter()
for _ in range(a.steps):
    step_session()
t = (time.perf_counter() - t0) / a.steps
e2e["e2e_session_two_cluster_calls"] = {"value": m / t, "unit": "reads/s", "ms_per_step": t * 1e3, "h2d_bytes_per_step": m * 20 + (n_jobs + 1) * 8,
                                        "d2h_bytes_per_step": m * 32}
same_session = bool(np.array_equal(h_rec2.numpy().view(pkg.UMI_CLUSTER_REC).reshape(m), rec_dev))     # ascending rank = no rank
same_e2e = same_session and bool(np.array_equal(h_rec.numpy().view(pkg.UMI_CLUSTER_REC).reshape(m), rec_dev))

# CPU oracle on a bounded sample of the same jobs (all host threads): matrices + clusterLocal's two passes
js = min(a.cpu_jobs, a.jobs)
ms_, os_ = int(offs[js]), oo[:js + 1]
t0 = time.perf_counter()
cm, _ = orc.umi_matrix_batch(umis[:ms_], offs[:js + 1], 12)
crec = orc.umi_cluster_batch(cm, offs[:js + 1], os_, a.ed)
tcpu = time.perf_counter() - t0
same = bool(crec.tobytes() == rec_dev[:ms_].tobytes() and np.array_equal(cm, d_m[:int(os_[-1])].cpu().numpy()))
alg = cells * 12 + m * 32                          # matrix written once (4 B/cell) and read by the two cluster passes (8 B/cell), codes in, records out
peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = float(json.load(open(peaks_path)).get("hbm_gbs", 6550.0)) if os.path.exists(peaks_path) else 6550.0
ach = alg / (ms / 1e3) / 1e9
pairs = int(((np.diff(offs) * (np.diff(offs) - 1)) // 2).sum())
from bench import issue_roofline
roof, _ = issue_roofline(pkg, "umi_pairs_kernel/pair", int(((np.diff(offs) * (np.diff(offs) - 1)) // 2).sum()), ms_dist, clocks.get("sm_mhz"),
                         torch.cuda.get_device_properties(0).multi_processor_count, unit="pair")
print(json.dumps({
    "metric": "reads/sec UMI distance matrices + neighbour-set clustering (ED %d)" % a.ed, "value": m / (ms / 1e3), "unit": "reads/s", "n_gpus": 1,
    "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "u32", "data": "synthetic",
    "config": {"workload": "umi_cluster: %d (cell, gene) jobs of geometric size (mean 4, cap 2000) + one job of %d reads = %d reads, %d matrix cells "
                           "(%.1f GB, larger than L2), 12-nt UMIs +-1" % (a.jobs, a.deep, m, cells, cells * 4 / 1e9)},
    "clocks": clocks, "gpu_launches": int(launches), "ms_distance_kernels": ms_dist, "ms_cluster_kernels": ms - ms_dist,
    "pairs_per_s": pairs / (ms_dist / 1e3), "keys_fraction": float((rec_dev["best_key"] >= 0).mean()),
    "e2e": e2e["e2e"], "e2e_records_only": e2e["e2e_records_only"], "e2e_session_two_cluster_calls": e2e["e2e_session_two_cluster_calls"],
    "e2e_matches_device": same_e2e,
    "cpu_baseline": {"value": ms_ / tcpu, "unit": "reads/s", "cores": os.cpu_count(), "kind": "port",
                     "sample": "first %d jobs (%d reads) of the batch, CPU oracle (orc_umi_matrix_batch + orc_umi_cluster_batch, OpenMP over jobs)" % (js, ms_),
                     "gpu_matches_oracle_on_sample": same},
    "roofline": roof,
    "roofline_hbm_cluster_kernels": {"bound": "hbm", "kernel": "umi_cluster_* (the two clusterLocal passes)", "achieved": (cells * 8 + m * 16) / ((ms - ms_dist) / 1e3) / 1e9,
                                     "peak": peak, "unit": "GB/s", "frac": (cells * 8 + m * 16) / ((ms - ms_dist) / 1e3) / 1e9 / peak, "traffic": None,
                                     "note": "matrix read by the two passes (8 B/cell) + records out"},
    "reference_equivalent_gbs": {"value": ach, "algorithmic_bytes_per_step": alg, "note": "matrix written once and read twice + codes in + records out, over the whole step"}}))
