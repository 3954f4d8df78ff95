"""Kernel timing of the Illumina-guided search (device-resident inputs, CUDA events).
   python tools/perf_guided.py <flavour umi|bc> <L> <ed> <pm> <n_queries> [group_size] [reps]"""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
flavour, L, ed, pm, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
gsize = int(sys.argv[6]) if len(sys.argv) > 6 else 8
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
bc = flavour == "bc"
post_len = 10 if bc else ed + pm + 2
w = pkg.synth_guided(n, L, seed=9, n_groups=max(1, n // 4), group_size=gsize, pm=pm, post_len=post_len, bc_flavour=bc)
ctx = pkg.Context(0)
sets = pkg.GuidedSets(ctx, w["group_keys"], w["group_offsets"], L, bc_flavour=bc, all_keys=w["all_keys"], all_ed=3, empty_keys=w["empty_keys"], empty_ed=2)
dev = torch.device("cuda", 0)
d_sl, d_an, d_gid = (torch.from_numpy(w[k]).to(dev) for k in ("slices", "anchor", "group_id"))
d_ed = torch.full((n,), ed, dtype=torch.int32, device=dev)
d_out = torch.empty((n, 40), dtype=torch.uint8, device=dev)
lib = pkg.gpu_lib()
st = torch.cuda.current_stream().cuda_stream
def run():
    pkg._check(lib.slr_guided_match_dev(ctx.h, sets.h, pm, post_len, 2 if bc else -1, d_sl.data_ptr(), 32, 32, d_an.data_ptr(), d_gid.data_ptr(),
                                        d_ed.data_ptr(), ed, n, d_out.data_ptr(), None, 0, st))
run(); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
res = d_out.cpu().numpy().view(pkg.GUIDED_RESULT).reshape(-1)
print(json.dumps({"flavour": flavour, "L": L, "ed": ed, "pm": pm, "queries": n, "group_size": gsize, "ms_best": min(ts), "queries_per_s": n / (min(ts) / 1e3),
                  "found": float((res["n_distinct"] > 0).mean()), "second": float((res["n_distinct"] > 1).mean()), "exceptions": int((res["flags"] != 0).sum()),
                  "crc": int(np.bitwise_xor.reduce(res["seq"][:, 0] * np.uint64(31) + res["n_raw"].astype(np.uint64)))}))
