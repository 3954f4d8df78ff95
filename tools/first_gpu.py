"""Quick timing probe used during development (not the bench): kernel-only and end-to-end times for a few shapes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
for (nwl, wseed, ed) in ((737280, 737, 1), (3_000_000, 3_000_000, 2), (737280, 737, 2), (3_000_000, 3_000_000, 1)):
    wl = pkg.synth_whitelist(nwl, wseed)
    t0 = time.time(); table = pkg.BarcodesMapForBCfinding.getMapFromCellRangerData(ctx, wl); t1 = time.time()
    sl, an, _ = pkg.synth_reads(wl, n, seed=2)
    d_sl = torch.from_numpy(sl).cuda(); d_an = torch.from_numpy(an).cuda()
    d_out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    p = pkg.Parser(ctx, table, ed)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): p.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): p.assign_barcodes_dev(d_sl.data_ptr(), 32, d_an.data_ptr(), n, d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    psl = torch.from_numpy(sl).pin_memory().numpy(); pan = torch.from_numpy(an).pin_memory().numpy()
    out = torch.empty((n, 32), dtype=torch.uint8).pin_memory().numpy().view(pkg.BC_RESULT).reshape(-1)
    p.assign_barcodes(psl, pan, out=out)
    t2 = time.time(); p.assign_barcodes(psl, pan, out=out); t3 = time.time()
    ass = (out["flags"] & 1).mean()
    print(f"wl={nwl} ed={ed} n={n}: table build {t1-t0:.2f}s kernel {ms:.2f} ms -> {n/ms/1e3:.2f} Mreads/s ; e2e {t3-t2:.3f}s -> {n/(t3-t2)/1e6:.2f} Mreads/s ; assigned {ass:.3f}", flush=True)
umis, offs = pkg.synth_umi_jobs(500000, mean=4.0, cap=2000, seed=4)
t0 = time.time(); m, oo = pkg.generate_distance_matrices(ctx, umis, offs); t1 = time.time()
t0 = time.time(); m, oo = pkg.generate_distance_matrices(ctx, umis, offs); t1 = time.time()
print(f"umi: {len(umis)} reads {len(m)} cells e2e {t1-t0:.3f}s -> {len(m)/(t1-t0)/1e6:.1f} Mcells/s")
print("launches", pkg.launch_count())
