"""Kernel-only timing of slr_bc_assign_dev (CUDA events), for kernel experiments:
   python tools/perf_bc.py <n_list> <list_seed> <ed> <n_reads> [reps]        (SLR_LIB_GPU=<variant .so> to pick a build)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
nwl, wseed, ed, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
ctx = pkg.Context(0)
wl = pkg.synth_whitelist(nwl, wseed)
table = pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, wl)
sl, an, _ = pkg.synth_reads(wl, n, seed=2)
d_sl = torch.from_numpy(sl).cuda(); d_an = torch.from_numpy(an).cuda()
d_out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
p = pkg.Parser(ctx, table, ed)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    p.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), st)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
res = d_out.cpu().numpy().view(pkg.BC_RESULT).reshape(-1)
import zlib
print(json.dumps({"lib": os.path.basename(pkg.LIB_GPU), "list": nwl, "ed": ed, "reads": n, "ms_best": min(ts), "ms_med": sorted(ts)[len(ts) // 2],
                  "Mreads_per_s": n / min(ts) / 1e3, "assigned": float((res["flags"] & 1).mean()), "crc": zlib.crc32(res.tobytes())}))
