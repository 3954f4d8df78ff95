"""Summarise one ncu --set full report into JSON (the numbers profiles/*.json and DESIGN.md quote).
   python tools/ncu_summary.py <report.ncu-rep> <units_per_launch> [out.json]"""
import csv, io, json, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row = rows[0], rows[1]
out = {"report": rep.split("/")[-1], "kernels": []}
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, unit_row))
    def f(k, scale=1.0):
        try:
            return float(d[k].replace(",", "")) * scale
        except Exception:
            return None
    def bytes_of(k):
        v = f(k)
        if v is None: return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u.get(k, "byte"), 1)
    def time_ms(k):
        v = f(k)
        if v is None: return None
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u.get(k, "ns"), 1e-6)
    stalls = sorted(((k.split("issue_stalled_")[1].split("_per_issue")[0], float(v)) for k, v in d.items()
                     if "average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v),
                    key=lambda x: -x[1])[:6]
    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    k = {"kernel": d.get("Kernel Name", "")[:90], "grid": d.get("Grid Size"), "block": d.get("Block Size"),
         "duration_ms": time_ms("gpu__time_duration.sum"), "units_per_launch": units,
         "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_unit": (rd + wr) / units if rd is not None else None,
         "dram_pct_of_peak": f("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
         "lts_sector_hit_rate_pct": f("lts__t_sector_hit_rate.pct"), "l1tex_sector_hit_rate_pct": f("l1tex__t_sector_hit_rate.pct"),
         "l2_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
         "global_load_sectors_per_unit": (f("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") or 0) / units,
         "local_store_requests_per_unit": (f("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum") or 0) / units,
         "inst_executed_per_unit": (f("smsp__inst_executed.sum") or 0) / units,
         "ipc_active": f("sm__inst_executed.avg.per_cycle_active"), "issue_slots_busy_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active"),
         "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
         "achieved_warps_per_sm": f("sm__warps_active.avg.per_cycle_active"), "registers": f("launch__registers_per_thread"),
         "avg_active_threads_per_warp": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
         "top_stalls_warps_per_issue": stalls}
    out["kernels"].append(k)
txt = json.dumps(out, indent=1)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt + "\n")
print(txt)
