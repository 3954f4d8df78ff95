"""Measurement of the Illumina-guided search row, same shape as bench.py's line: device-resident throughput (CUDA events),
the host-pointer C-ABI call with H2D / D2H inside the timed region (`e2e`), the CPU oracle on a bounded sample (`cpu_baseline`) and the
roofline bookkeeping.  One B200:   python tools/bench_guided.py [--flavour umi|bc] [--ed 2] [--reads 2000000] [--steps 5] [--warmup 3]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
from bench import ClockSampler, issue_roofline

ap = argparse.ArgumentParser()
ap.add_argument("--flavour", default="umi", choices=["umi", "bc"])
ap.add_argument("--ed", type=int, default=2)
ap.add_argument("--pm", type=int, default=2)
ap.add_argument("--reads", type=int, default=2_000_000)
ap.add_argument("--group-size", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu-sample", type=int, default=0)
a = ap.parse_args()
pkg = g.load_package()
from oracle import orc
bc = a.flavour == "bc"
L = 16 if bc else 12
post_len = 10 if bc else a.ed + a.pm + 2
gsize = a.group_size or (300 if bc else 8)
bailout = 2 if bc else None
n = a.reads
w = pkg.synth_guided(n, L, seed=9, n_groups=max(1, n // (gsize * 2 if bc else 4)), group_size=gsize, pm=a.pm, post_len=post_len, bc_flavour=bc)
ctx = pkg.Context(0)
sets = pkg.GuidedSets(ctx, w["group_keys"], w["group_offsets"], L, bc_flavour=bc, all_keys=w["all_keys"], all_ed=3, empty_keys=w["empty_keys"], empty_ed=2)
dev = torch.device("cuda", 0)
pin = lambda x: torch.from_numpy(x).pin_memory()
h_sl, h_an, h_gid = pin(w["slices"]), pin(w["anchor"]), pin(w["group_id"])
h_ed = pin(np.full(n, a.ed, dtype=np.int32))
h_out = torch.empty((n, 40), dtype=torch.uint8).pin_memory()
d_sl, d_an, d_gid, d_ed = (t.to(dev) for t in (h_sl, h_an, h_gid, h_ed))
d_out = torch.empty((n, 40), dtype=torch.uint8, device=dev)
lib, st = pkg.gpu_lib(), torch.cuda.current_stream().cuda_stream

def step_device():
    pkg._check(lib.slr_guided_match_dev(ctx.h, sets.h, a.pm, post_len, -1 if bailout is None else bailout, d_sl.data_ptr(), 32, 32, d_an.data_ptr(),
                                        d_gid.data_ptr(), d_ed.data_ptr(), a.ed, n, d_out.data_ptr(), None, 0, st))

def step_e2e():
    pkg._check(lib.slr_guided_match(ctx.h, sets.h, a.pm, post_len, -1 if bailout is None else bailout, h_sl.data_ptr(), 32, 32, h_an.data_ptr(),
                                    h_gid.data_ptr(), h_ed.data_ptr(), n, h_out.data_ptr(), None, 0))

for _ in range(a.warmup):
    step_device()
torch.cuda.synchronize()
l0 = pkg.launch_count()
sampler = ClockSampler(0)
sampler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step_device()
e1.record()
torch.cuda.synchronize()
clocks = sampler.summary()
sampler.join(timeout=10)
launches = pkg.launch_count() - l0
ms = e0.elapsed_time(e1) / a.steps
for _ in range(max(2, a.warmup)):
    step_e2e()
t0 = time.perf_counter()
for _ in range(a.steps):
    step_e2e()
e2e_ms = (time.perf_counter() - t0) / a.steps * 1e3
res = d_out.cpu().numpy().view(pkg.GUIDED_RESULT).reshape(-1)
# CPU oracle on a bounded sample (all host threads) + the reference's probe count per read
n_s = a.cpu_sample or {0: 2_000_000, 1: 1_000_000, 2: 20_000, 3: 2_000, 4: 600}[a.ed] // (6 if bc else 1)
n_s = min(n, max(n_s, 64))
kw = dict(bailout=-1 if bailout is None else bailout, bc_flavour=bc, all_keys=w["all_keys"], all_ed=3, empty_keys=w["empty_keys"], empty_ed=2)
t0 = time.perf_counter()
cres, _, probes = orc.guided_batch(w["group_keys"], w["group_offsets"], w["slices"][:n_s], w["anchor"][:n_s], w["group_id"][:n_s], a.ed, L, a.pm, post_len, **kw)
tcpu = time.perf_counter() - t0
same = bool((cres == res[:n_s]).all())
ppr = probes / n_s
alg = ppr * 8 + 32 + 12 + 40                                    # SURVEY 8d convention: probes x 8 B + boundary in (slice, anchor, group, ed) / out
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6550.0))
ach = n * alg / (ms / 1e3) / 1e9
roof, _ = issue_roofline(pkg, "guided_match_kernel<%s,%d>" % ("bc" if bc else "umi", a.ed), n, ms, clocks.get("sm_mhz"),
                         torch.cuda.get_device_properties(0).multi_processor_count)
print(json.dumps({
    "metric": "reads/sec Illumina-guided %s search (ED %d, +-%d)" % ("cell barcode" if bc else "UMI", a.ed, a.pm), "value": n / (ms / 1e3), "unit": "reads/s",
    "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "u32", "data": "synthetic",
    "config": {"workload": "guided_%s_ed%d: %d synthetic stranded slices, L %d, candidate groups of ~%d%s, bailout %s" % (
        a.flavour, a.ed, n, L, gsize, " + %d all-passed + %d empty-drop barcodes" % (len(w["all_keys"]), len(w["empty_keys"])) if bc else "", bailout)},
    "clocks": clocks, "gpu_launches": int(launches), "found_fraction": float((res["n_distinct"] > 0).mean()), "second_fraction": float((res["n_distinct"] > 1).mean()),
    "e2e": {"value": n / (e2e_ms / 1e3), "unit": "reads/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 44, "d2h_bytes_per_step": n * 40},
    "cpu_baseline": {"value": n_s / tcpu, "unit": "reads/s", "cores": os.cpu_count(), "kind": "port", "sample": "first %d reads of the batch, CPU oracle (orc_guided_batch, OpenMP over reads)" % n_s,
                     "probes_per_read": ppr, "gpu_matches_oracle_on_sample": same},
    "roofline": roof,
    "reference_equivalent_gbs": {"value": ach, "algorithmic_bytes_per_read": alg, "note": "bytes the REFERENCE algorithm would touch (hash probes x 8 B + boundary in/out) "
                                 "divided by the kernel time: a work-equivalence figure, not a utilisation (HBM peak %.0f GB/s)" % peak}}))
