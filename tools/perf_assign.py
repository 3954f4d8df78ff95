"""Kernel-only timing of slr_umi_assign_dev2 (ClusterOneHierarchical / ClusterOne_MyClustering on the resident matrices) beside slr_umi_dist_dev, with the CPU oracle
for parity + baseline: python tools/perf_assign.py <n_reads> [mean] [cap] [reps]"""
import sys, os, json, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]); mean = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = pkg.Context(0)
umis, offs = pkg.synth_umi_jobs(int(n / mean * 1.05) + 1000, mean=mean, cap=cap, seed=4)
k = int(np.searchsorted(offs, n, side="right")) - 1
offs = offs[:k + 1].copy(); umis = np.ascontiguousarray(umis[:offs[-1]])
oo = pkg.out_offsets_for(offs); cells = int(oo[-1]); m = len(umis); nj = len(offs) - 1
d_u, d_o, d_oo = (torch.from_numpy(x).cuda() for x in (umis, offs, oo))
d_m = torch.empty(cells, dtype=torch.int32, device="cuda")
d_rec = torch.empty((m, 16), dtype=torch.uint8, device="cuda")
lib = pkg.gpu_lib(); st = torch.cuda.current_stream().cuda_stream
scr_bytes = int(lib.slr_umi_assign_scratch_bytes(nj)) + sum(int(lib.slr_umi_assign_deep_job_bytes(int(n))) for n in np.diff(offs) if n > 100)
d_scr = torch.empty(scr_bytes, dtype=torch.uint8, device="cuda")
def dist():
    pkg._check(lib.slr_umi_dist_dev(ctx.h, d_u.data_ptr(), 16, 12, d_o.data_ptr(), nj, m, d_m.data_ptr(), d_oo.data_ptr(), cells, st))
def assign():
    pkg._check(lib.slr_umi_assign_dev2(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), nj, m, None, None, d_scr.data_ptr(), scr_bytes, d_rec.data_ptr(), st))
def timed(f):
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
t_dist = timed(dist); t_as = timed(assign)
rec = d_rec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(m)
sizes = np.diff(offs)
out = {"lib": os.path.basename(pkg.LIB_GPU), "reads": m, "jobs": nj, "jobs_ge2": int((sizes >= 2).sum()), "jobs_gt32": int((sizes > 32).sum()),
       "jobs_gt100": int((sizes > 100).sum()), "cells": cells, "dist_ms": t_dist, "assign_ms": t_as, "assign_Mreads_per_s": m / t_as / 1e3,
       "assign_Mjobs_per_s": nj / t_as / 1e3, "assigned": int((rec["flags"] & 1 != 0).sum()), "tie_unpin_reads": int((rec["flags"] & 4 != 0).sum()),
       "deep_reads": int((rec["flags"] & 8 != 0).sum()), "deep_assigned": int(((rec["flags"] & 9) == 9).sum()), "crc": zlib.crc32(rec.tobytes())}
if cells <= 3e8:
    from oracle import orc
    orc.build()
    mats = d_m.cpu().numpy()
    t0 = time.perf_counter(); want = orc.umi_assign_batch(mats, offs, oo); dt = time.perf_counter() - t0
    out.update({"oracle_ms": dt * 1e3, "oracle_threads": os.cpu_count(), "parity": bool(want.tobytes() == rec.tobytes())})
print(json.dumps(out))
