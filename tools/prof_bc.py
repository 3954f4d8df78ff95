"""One configuration, few launches: the target of ncu captures (see profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
nwl, wseed, ed, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = pkg.Context(0)
wl = pkg.synth_whitelist(nwl, wseed)
table = pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, wl)
sl, an, _ = pkg.synth_reads(wl, n, seed=2)
d_sl = torch.from_numpy(sl).cuda(); d_an = torch.from_numpy(an).cuda()
d_out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
p = pkg.Parser(ctx, table, ed)
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    p.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), st)
torch.cuda.synchronize()
print("done", (d_out.cpu().numpy().view(pkg.BC_RESULT)["flags"] & 1).mean())
