"""Kernel-only timing of the large-job path (umi_assign_deep.cu) per job size: python tools/perf_deep.py [sizes ...]
each size N is timed as ONE job of N reads (synthetic deep job, as bench.py's umi5kx2k); the last line is a batch of 256 jobs of ~300 reads"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0)
lib = pkg.gpu_lib(); st = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [300, 1000, 2000, 4000, 8000, 20000]


def run(umis, offs, label):
    oo = pkg.out_offsets_for(offs); cells = int(oo[-1]); m = len(umis); nj = len(offs) - 1
    d_u, d_o, d_oo = (torch.from_numpy(x).cuda() for x in (umis, offs, oo))
    d_m = torch.empty(cells, dtype=torch.int32, device="cuda")
    d_rec = torch.empty((m, 16), dtype=torch.uint8, device="cuda")
    nb = int(lib.slr_umi_assign_scratch_bytes(nj)) + sum(int(lib.slr_umi_assign_deep_job_bytes(int(n))) for n in np.diff(offs) if n > 100)
    d_scr = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pkg._check(lib.slr_umi_dist_dev(ctx.h, d_u.data_ptr(), 16, 12, d_o.data_ptr(), nj, m, d_m.data_ptr(), d_oo.data_ptr(), cells, st))
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg._check(lib.slr_umi_assign_dev2(ctx.h, d_m.data_ptr(), d_o.data_ptr(), d_oo.data_ptr(), nj, m, None, None, d_scr.data_ptr(), nb, d_rec.data_ptr(), st))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    rec = d_rec.cpu().numpy().view(pkg.UMI_ASSIGN_REC).reshape(m)
    if nj == 1:                                                # phase stamps of the job (header words 32..63 of its arena, umi_assign_deep.cu)
        base = int(lib.slr_umi_assign_scratch_bytes(1))
        # arena offset inside the scratch: 32 + (4 n_jobs + 4) * 4 + (2 n_jobs + 2) * 8 bytes
        arena = 32 + (4 * nj + 4) * 4 + (2 * nj + 2) * 8
        hdr = d_scr[arena + 2 * ((m + 3) & ~1) * 4: arena + 2 * ((m + 3) & ~1) * 4 + 256].cpu().numpy().view(np.uint64)[16:32].astype(np.int64)
        names = ["r1 counts", "r1 keys", "r1 choose", "r1 scan", "r1 layout", "r1 build", "r1 centres", "(victims/remove)", "r2 counts", "r2 keys", "r2 choose",
                 "r2 scan", "r2 layout", "r2 build", "r2 centres+per-cluster", "per-read"]
        seq = [(i, int(hdr[i])) for i in range(16) if hdr[i] > 0]
        print("   phases (us):", ", ".join("%s %.0f" % (names[a], (tb - ta) / 1e3) for (a, ta), (b, tb) in zip(seq, seq[1:])), flush=True)
    parity = None
    if m <= 30000:                                             # the CPU oracle on the same matrices (its O(n^2) passes use all host threads)
        from oracle import orc
        want = orc.umi_assign_batch(d_m.cpu().numpy(), offs, oo)
        parity = bool(want.tobytes() == rec.tobytes())
    print(json.dumps({"what": label, "reads": m, "jobs": nj, "parity": parity, "assign_ms": min(ts), "assigned": int((rec["flags"] & 1 != 0).sum()),
                      "clusters": int(rec["n_clusters"][0]), "tie_unpin": bool(rec["flags"][0] & 4), "scratch_MB": nb / 1e6}), flush=True)


for N in sizes:
    u, o = pkg.synth_umi_jobs(1, mean=1e9, cap=N, seed=77)
    run(u, o, "one job of %d reads" % N)
u, o = pkg.synth_umi_jobs(256, mean=300.0, cap=900, seed=5)
run(u, o, "256 jobs of ~300 reads")
