"""Kernel-only timing of slr_umi_cluster_dev on the matrices slr_umi_dist_dev leaves on the device (CUDA events), with the
16-thread CPU oracle beside it: python tools/perf_cluster.py <n_reads> [mean] [cap] [ed] [reps]"""
import sys, os, json, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]); mean = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
ed = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
ctx = pkg.Context(0)
umis, offs = pkg.synth_umi_jobs(int(n / mean * 1.05) + 1000 if mean < 1e6 else 1, mean=mean, cap=cap, seed=4)
if mean < 1e6:
    k = int(np.searchsorted(offs, n, side="right")) - 1
    offs = offs[:k + 1].copy(); umis = np.ascontiguousarray(umis[:offs[-1]])
oo = pkg.out_offsets_for(offs); cells = int(oo[-1]); m = len(umis)
d_u, d_o, d_oo = (torch.from_numpy(x).cuda() for x in (umis, offs, oo))
d_m = torch.empty(cells, dtype=torch.int32, device="cuda")
d_cnt = torch.empty(m, dtype=torch.int32, device="cuda"); d_rec = torch.empty(m * 4, dtype=torch.int32, device="cuda")
lib = pkg.gpu_lib(); st = torch.cuda.current_stream().cuda_stream
pkg._check(lib.slr_umi_dist_dev(ctx.h, d_u.data_ptr(), 16, 12, d_o.data_ptr(), len(offs) - 1, m, d_m.data_ptr(), d_oo.data_ptr(), cells, st))
def run():
    pkg._check(lib.slr_umi_cluster_dev(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), len(offs) - 1, m, ed, None, None,
                                       d_cnt.data_ptr(), d_rec.data_ptr(), st))
for _ in range(2): run()
torch.cuda.synchronize(); ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
rec = d_rec.cpu().numpy().view(pkg.UMI_CLUSTER_REC).reshape(m)
out = {"lib": os.path.basename(pkg.LIB_GPU), "reads": m, "jobs": len(offs) - 1, "cells": cells, "ed": ed, "ms_best": min(ts),
       "Mreads_per_s": m / min(ts) / 1e3, "matrix_GBps": 2 * cells * 4 / min(ts) / 1e6, "keys": int((rec["best_key"] >= 0).sum()),
       "tied": int((rec["n_ties"] > 1).sum()), "crc": zlib.crc32(rec.tobytes())}
if cells <= 3e8:                                      # CPU oracle on the same matrices (all host threads), parity + baseline
    from oracle import orc
    orc.build()
    mats = d_m.cpu().numpy()
    t0 = time.perf_counter(); want = orc.umi_cluster_batch(mats, offs, oo, ed); dt = time.perf_counter() - t0
    out.update({"oracle_ms": dt * 1e3, "oracle_threads": os.cpu_count(), "parity": bool(want.tobytes() == rec.tobytes())})
print(json.dumps(out))
