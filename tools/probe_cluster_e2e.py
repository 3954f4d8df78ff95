import sys, os, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0, 2)
n = 10_000_000
umis, offs = pkg.synth_umi_jobs(int(n / 4 * 1.05) + 1000, mean=4.0, cap=2000, seed=4)
k = int(np.searchsorted(offs, n, side="right")) - 1
offs = offs[:k + 1].copy(); umis = np.ascontiguousarray(umis[:offs[-1]])
res = {}
for name, fn in (("umi_dist (matrices back)", lambda: pkg.generate_distance_matrices(ctx, umis, offs)),
                 ("umi_cluster (records only)", lambda: pkg.cluster_local(ctx, umis, offs, 2)),
                 ("umi_cluster (records + matrices)", lambda: pkg.cluster_local(ctx, umis, offs, 2, want_matrices=True))):
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    res[name] = round(min(ts[1:]), 2)
print(json.dumps({"reads": len(umis), "host_call_ms": res}))
