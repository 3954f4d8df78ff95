"""All GPUs of the box from ONE process through slr_multi_* (the route a single JVM takes): end-to-end reads/s of the bc3m_ed2 step
(slr_multi_bc_assign + slr_multi_umi_assign from two host threads, host buffers, copies inside), parity against the oracle on a strided
sample, merged counters.  python tools/bench_multi.py [n_devices] [reads_per_gpu] [steps]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
import __graft_entry__ as g
pkg = g.load_package()
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mg = pkg.MultiGpu(nd)
nd = mg.n_devices
n = nd * R
wl = pkg.synth_whitelist(3_000_000, 3_000_000)
t0 = time.perf_counter(); mg.load_barcodes(wl, np.arange(1, len(wl) + 1, dtype=np.int32)); t_table = time.perf_counter() - t0
pin = lambda t: t.pin_memory()
h_sl, h_an, h_res = pin(torch.empty((n, 32), dtype=torch.uint8)), pin(torch.empty(n, dtype=torch.int32)), pin(torch.empty((n, 32), dtype=torch.uint8))
pkg.synth_reads(wl, n, seed=2, out=(h_sl.numpy(), h_an.numpy()))
umis, offs, _, _ = pkg.synth_umi_shard(0, n, 4.0, 2000, 4)
h_um, h_of = pin(torch.from_numpy(umis)), pin(torch.from_numpy(offs))
h_ar = pin(torch.empty((n, 16), dtype=torch.uint8))
res = h_res.numpy().view(pkg.BC_RESULT).reshape(-1); arec = h_ar.numpy().view(pkg.UMI_ASSIGN_REC).reshape(-1)
pool = ThreadPoolExecutor(2)
def step():
    f = pool.submit(mg.assign_barcodes, h_sl.numpy(), h_an.numpy(), 2, 2, True, None, res)
    mg.umi_assign(h_um.numpy(), h_of.numpy(), out=arec)
    f.result()
for _ in range(3): step()
mg.reset_counts()
ts = []
for _ in range(steps):
    t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
counts = mg.counts()
ok = (res["flags"] & 1) != 0
from oracle import orc
sel = np.arange(0, n, max(1, n // 4000))
exp, _ = orc.assign_barcode_batch(orc.BarcodeSet(wl, np.arange(1, len(wl) + 1, dtype=np.int32)), h_sl.numpy()[sel], h_an.numpy()[sel], 2)
jsel = np.arange(0, len(offs) - 1, max(1, (len(offs) - 1) // 20000))
su = np.concatenate([umis[offs[j]:offs[j + 1]] for j in jsel]); so = np.concatenate([[0], np.cumsum([offs[j + 1] - offs[j] for j in jsel])]).astype(np.int64)
em, eo = orc.umi_matrix_batch(su, so); er = orc.umi_assign_batch(em, so, eo)
umi_ok = all(arec[offs[j]:offs[j + 1]].tobytes() == er[so[k]:so[k + 1]].tobytes() for k, j in enumerate(jsel))
ms = sum(ts) / len(ts)
print(json.dumps({"tool": "bench_multi", "n_gpus": nd, "single_process": True, "peer_access": mg.peer_access, "reads_per_gpu": R,
                  "e2e_reads_per_s": n / ms * 1e3, "ms_per_step": ms, "ms_steps": [round(x, 2) for x in ts], "table_build_all_devices_ms": t_table * 1e3,
                  "h2d_bytes_per_step": n * 52 + 8 * len(offs), "d2h_bytes_per_step": n * 48,
                  "parity": {"bc_sample": bool((res[sel] == exp).all()), "umi_sample": bool(umi_ok),
                             "counters": bool(counts.sum() == steps * ok.sum() and (counts.sum(axis=0) == steps * np.bincount(res["ed"][ok], minlength=3)).all())}}))
