"""Timing of slr_bc_collide (host-pointer call, CUDA kernel inside) vs the CPU oracle: python tools/perf_collide.py <n_used> [ed]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import orc
n = int(sys.argv[1]); ed = int(sys.argv[2]) if len(sys.argv) > 2 else 2
base = pkg.synth_whitelist(n * 6 // 10, 11)
rng = np.random.default_rng(5)
kids = base[rng.integers(0, len(base), n - len(base))] ^ (np.uint64(1) << rng.integers(0, 32, n - len(base)).astype(np.uint64))
wl = np.unique(np.concatenate([base, kids]))
ctx = pkg.Context(0)
table = pkg.BarcodesMapForBCfinding(ctx, wl)
t = pkg.BarcodeDatasetColissionTester(ctx, table, ed)
t.colissionsFromScan()
t0 = time.perf_counter(); got = t.colissionsFromScan(); t1 = time.perf_counter()
ns = min(len(wl), 20000)
c0 = time.perf_counter(); exp, probes = orc.collide_batch(orc.BarcodeSet(wl), wl[:ns], ed); c1 = time.perf_counter()
print(json.dumps({"used_barcodes": len(wl), "ed": ed, "gpu_ms_e2e": (t1 - t0) * 1e3, "gpu_barcodes_per_s": len(wl) / (t1 - t0),
                  "cpu_oracle_barcodes_per_s": ns / (c1 - c0), "cpu_threads": os.cpu_count(), "ref_probes_per_barcode": probes / ns,
                  "match_on_sample": bool((got[:ns] == exp).all()), "with_ed1_neighbour": float((got["valid"] & 1).mean())}))
