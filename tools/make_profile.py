"""Per-kernel profile the bench's roofline reads (run on the GPU box):  python tools/make_profile.py [out.json]

One ncu pass (--clock-control none, a few counters, no --set full) over short single-purpose runs of the hot kernels; per kernel the warp
instructions, DRAM bytes and duration of one launch, divided by the units (reads) the launch processed.  The file carries the sha256 of the
library's SOURCES (pkg.csrc_sha256(): csrc/ + the public header) and, per kernel, of that kernel's translation unit + transitive includes
(pkg.kernel_src_sha256): bench.py uses an entry only when the kernel's hash equals the one of the sources the loaded library was built from,
i.e. the numbers describe the code that ran (the .so itself is not bit-reproducible across rebuilds: nvcc's anonymous-namespace names)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_kernel_profile.json")
METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,"
           "smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct,sm__warps_active.avg.per_cycle_active,"
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread")
RUNS = [
    # (profile key, kernel regex, units per launch, command, what)
    ("bc_assign_kernel<2>", "bc_assign_kernel", 1_000_000, ["tools/prof_bc.py", "3000000", "3000000", "2", "1000000", "1"], "1 M reads, 3 M list, ED 2"),
    ("bc_assign_kernel<1>", "bc_assign_kernel", 4_000_000, ["tools/prof_bc.py", "737280", "737", "1", "4000000", "1"], "4 M reads, 737 K list, ED 1"),
    ("umi_pairs_kernel", "umi_pairs_kernel", 4_000_000, ["tools/perf_assign.py", "4000000", "4", "2000", "1"], "4 M reads in jobs of mean 4"),
    ("umi_assign_kernel", "umi_assign_kernel", 4_000_000, ["tools/perf_assign.py", "4000000", "4", "2000", "1"], "4 M reads in jobs of mean 4"),
    # the side benches: units = what their own JSON line reports (roofline.units_per_launch)
    # the same kernel per READ PAIR (units = pairs of the run, from the script's own JSON line): what tools/bench_umi_cluster.py scales by
    ("umi_pairs_kernel/pair", "umi_pairs_kernel", "pairs", ["tools/perf_assign.py", "2000000", "12", "2000", "1"], "2 M reads in jobs of mean 12, per read pair"),
    ("guided_match_kernel<umi,2>", "guided_match_kernel", None, ["tools/bench_guided.py", "--flavour", "umi", "--ed", "2", "--steps", "1", "--warmup", "1", "--cpu-sample", "64"],
     "tools/bench_guided.py defaults, UMI flavour, ED 2"),
    ("guided_match_kernel<bc,2>", "guided_match_kernel", None, ["tools/bench_guided.py", "--flavour", "bc", "--ed", "2", "--steps", "1", "--warmup", "1", "--cpu-sample", "64"],
     "tools/bench_guided.py defaults, BC flavour, ED 2"),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3,
        "msecond": 1.0, "second": 1e3}
kernels = []
for key, rx, units, cmd, what in RUNS:
    try:
        r = subprocess.run(["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:" + rx, "--csv", sys.executable] + cmd,
                           capture_output=True, text=True, cwd=ROOT, timeout=420)
    except subprocess.TimeoutExpired:
        print("timeout:", key, cmd, file=sys.stderr)
        continue
    lines = [l for l in r.stdout.splitlines() if l.startswith('"')]
    if units == "pairs":
        units = None
        for l in r.stdout.splitlines():
            if l.startswith("{") and '"cells"' in l:
                d = json.loads(l)
                units = (int(d["cells"]) - int(d["reads"])) // 2
        if not units:
            print("no JSON line from", cmd, r.stdout[-300:], r.stderr[-300:], file=sys.stderr)
            continue
    if units is None:
        for l in r.stdout.splitlines():
            if l.startswith("{") and "units_per_launch" in l:
                units = int(json.loads(l)["roofline"]["units_per_launch"])
        if units is None:
            print("no JSON line from", cmd, r.stdout[-300:], r.stderr[-300:], file=sys.stderr)
            continue
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    per = {}
    for row in rows:                                       # long format: one row per (launch id, metric)
        per.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row.get("Grid Size"), "block": row.get("Block Size")})
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per[row["ID"]][row["Metric Name"]] = v * UNIT.get(row["Metric Unit"], 1)
    if not per:
        print("no launch of", rx, "captured:", r.stdout[-500:], r.stderr[-500:], file=sys.stderr)
        continue
    # one pass of the hot path = one launch of every instance of the kernel (the large-job variant of a templated kernel is a separate
    # launch); the script behind `cmd` repeats the pass (warm-up + timed), so launches are averaged per instance and the instances summed
    best = max(per.values(), key=lambda d: d.get("gpu__time_duration.sum", 0))
    by_name = {}
    for d in per.values():
        by_name.setdefault(d["name"], []).append(d)
    tot = {k: sum(sum(d.get(k, 0) for d in ds) / len(ds) for ds in by_name.values())
           for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    kernels.append({"kernel": key, "src_sha256": pkg.kernel_src_sha256(key), "launch": best["name"][:100], "what": what, "instances": len(by_name), "launches_seen": len(per), "units_per_launch": units,
                    "duration_ms": tot["gpu__time_duration.sum"], "inst_executed_per_unit": tot["smsp__inst_executed.sum"] / units,
                    "dram_bytes_per_unit": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / units,
                    "issue_slots_busy_pct": best.get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                    "alu_pipe_pct": best.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "avg_active_threads_per_warp": best.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    "lts_sector_hit_rate_pct": best.get("lts__t_sector_hit_rate.pct"), "achieved_warps_per_sm": best.get("sm__warps_active.avg.per_cycle_active"),
                    "global_load_sectors_per_unit": best.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0) / units,
                    "registers": best.get("launch__registers_per_thread"), "grid": best.get("grid"), "block": best.get("block")})
    json.dump({"csrc_sha256": pkg.csrc_sha256(), "lib_sha256": pkg.lib_sha256(), "how": "ncu --metrics ... --clock-control none (tools/make_profile.py), durations are "
               "cold-cache profiler times: the bench uses the per-unit COUNTS only and its own CUDA-event times", "kernels": kernels},
              open(out_path, "w"), indent=1)                # rewritten after every kernel: a slow side bench cannot lose the earlier ones
print(open(out_path).read())
