"""Breakdown of the end-to-end (host-pointer) step: bc alone, umi alone, both concurrently; several repetitions each."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
import __graft_entry__ as g
import bench
pkg = g.load_package()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = pkg.Context(0, n_streams=2)
wl = pkg.synth_whitelist(3_000_000, 3_000_000)
table = pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, wl)
parser = pkg.Parser(ctx, table, 2)
pin = lambda t: t.pin_memory()
h_sl = pin(torch.empty((R, 32), dtype=torch.uint8)); h_an = pin(torch.empty(R, dtype=torch.int32))
pkg.synth_reads(wl, R, seed=2, out=(h_sl.numpy(), h_an.numpy()))
h_res = pin(torch.empty((R, 32), dtype=torch.uint8)); np_res = h_res.numpy().view(pkg.BC_RESULT).reshape(-1)
um, of = bench.umi_jobs_for(pkg, R, 4); oo = pkg.out_offsets_for(of)
h_um, h_of, h_oo = pin(torch.from_numpy(um)), pin(torch.from_numpy(of)), pin(torch.from_numpy(oo))
h_mat = pin(torch.empty(int(oo[-1]), dtype=torch.int32))
def bc(): parser.assign_barcodes(h_sl.numpy(), h_an.numpy(), None, np_res)
def umi(): pkg.generate_distance_matrices(ctx, h_um.numpy(), h_of.numpy(), 12, out=h_mat.numpy(), out_offsets=h_oo.numpy())
pool = ThreadPoolExecutor(max_workers=2)
def both():
    f = pool.submit(bc); umi(); f.result()
for name, fn in (("bc", bc), ("umi", umi), ("both", both), ("bc", bc), ("both", both)):
    fn(); ts = []
    for _ in range(6):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, " ".join("%.1f" % t for t in ts), flush=True)
