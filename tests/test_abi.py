"""The C-ABI library loads without a GPU and exports exactly what include/*.h declare; compute calls fail
loudly (SLR_E_NODEVICE) instead of falling back to the CPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    """every function the public headers declare: include/sicelore_gpu.h (the GPU seams) + include/sicelore_host.h (the host-side entry points)"""
    import glob
    src = "".join(open(f).read() for f in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    lib = pkg.gpu_lib()
    names = declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(pkg.EXPORTS) == names
    assert lib.slr_abi_version() == 1


def test_result_struct_layout(pkg, orc):
    assert pkg.BC_RESULT.itemsize == 32 and pkg.BC_RESULT == orc.BC_RESULT
    assert [pkg.BC_RESULT.fields[f][1] for f in ("bc", "ed", "ed_second", "offset", "n_ins", "n_del", "n_sub", "rank", "flags")] == \
        [0, 8, 12, 16, 17, 18, 19, 20, 24]


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.SiceloreGpuError) as e:
        pkg.Context(0)
    assert e.value.code == pkg.SLR_E_NODEVICE and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """the product path must never import / link the oracle"""
    bad = []
    pk = os.path.join(ROOT, "sicelore-2.1_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"(from|import)\s+oracle|slr_oracle\.h|liborc|orc_[a-z_]+\(", txt):
                    bad.append(f)
    assert not bad, bad


def test_synth_is_deterministic_and_shardable(pkg):
    import numpy as np
    wl = pkg.synth_whitelist(5000, 9)
    assert len(np.unique(wl)) == 5000 and (wl >> 32 == 0).all()
    a = pkg.synth_reads(wl, 1000, seed=5)
    b0 = pkg.synth_reads(wl, 400, seed=5)
    b1 = pkg.synth_reads(wl, 600, seed=5, first=400)
    assert (a[0] == np.concatenate([b0[0], b1[0]])).all() and (a[2] == np.concatenate([b0[2], b1[2]])).all()
    assert (a[1] == 8).all()
    u, o = pkg.synth_umi_jobs(50, mean=4.0, cap=30, seed=2)
    u2, o2 = pkg.synth_umi_jobs(50, mean=4.0, cap=30, seed=2)
    assert (u == u2).all() and (o == o2).all() and o[-1] == len(u) and (np.diff(o) >= 1).all()


# ---- callers other than Python: the JNI glue (compile / link check against a stub jni.h) and a plain-C driver ---------------
def build_c(out, src, extra):
    import subprocess
    lib_dir = os.path.join(ROOT, "sicelore-2.1_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include")] + extra +
                          [src, "-L" + lib_dir, "-lsicelore_gpu", "-Wl,-rpath," + lib_dir, "-o", out])
    return out


def test_jni_glue_compiles_and_links(pkg, tmp_path):
    """java/sicelore_gpu_jni.c forwards every native method of java/com/rw/gpu/Native.java to the C ABI"""
    import subprocess
    so = build_c(str(tmp_path / "libsicelore_gpu_jni.so"), os.path.join(ROOT, "java", "sicelore_gpu_jni.c"),
                 ["-shared", "-fPIC", "-I" + os.path.join(ROOT, "tests", "jni_stub")])
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r"Java_com_rw_gpu_Native_(\w+)", syms))
    java = open(os.path.join(ROOT, "java", "com", "rw", "gpu", "Native.java")).read()
    natives = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    assert natives and exported == natives, (sorted(natives - exported), sorted(exported - natives))


def write_driver_file(path, kind, arrays):
    import numpy as np
    with open(path, "wb") as f:
        f.write(b"SLRB" + np.uint32(kind).tobytes())
        for a in arrays:
            f.write(np.int64(a).tobytes() if np.isscalar(a) or isinstance(a, (int, np.integer)) else np.ascontiguousarray(a).tobytes())


def pkg_synth_assign_batch():
    import numpy as np
    """400 jobs of mean 6 and three jobs above 100 reads (the synthetic UMI generator needs no GPU)"""
    import __graft_entry__ as g
    pkg = g.load_package()
    u1, o1 = pkg.synth_umi_jobs(400, mean=6.0, cap=90, seed=3)
    u2, o2 = pkg.synth_umi_jobs(3, mean=1e9, cap=180, seed=9)
    return np.concatenate([u1, u2]), np.concatenate([o1, o1[-1] + o2[1:]]).astype(np.int64)


def driver_inputs(tmp_path, orc):
    import numpy as np
    gold = os.path.join(ROOT, "tests", "golden")
    files = []
    g = np.load(os.path.join(gold, "bc_3p_ed2.npz"))
    n = len(g["slices"])
    files.append(write_driver_file(tmp_path / "bc.bin", 1, [len(g["whitelist"]), 1, g["whitelist"], g["rank"], int(g["ed"]), 2, int(g["three_prime"]),
                                                          n, g["slices"], g["anchor"], g["result"]]) or tmp_path / "bc.bin")
    ex = orc.exact_lookup_batch(orc.BarcodeSet(g["whitelist"], g["rank"]), g["slices"], g["anchor"], True)
    files.append(write_driver_file(tmp_path / "exact.bin", 4, [len(g["whitelist"]), 1, g["whitelist"], g["rank"], 0, 0, 1, n, g["slices"],
                                                             g["anchor"], ex]) or tmp_path / "exact.bin")
    c = np.load(os.path.join(gold, "collide_ed2_skew.npz"))
    files.append(write_driver_file(tmp_path / "collide.bin", 3, [len(c["whitelist"]), 0, c["whitelist"], int(c["ed"]), len(c["whitelist"]),
                                                               c["whitelist"], c["result"]]) or tmp_path / "collide.bin")
    u = np.load(os.path.join(gold, "umi_len12.npz"))
    files.append(write_driver_file(tmp_path / "umi.bin", 2, [int(u["umi_len"]), len(u["job_offsets"]) - 1, len(u["umis"]), len(u["matrix"]),
                                                           u["umis"], u["job_offsets"], u["out_offsets"], u["matrix"]]) or tmp_path / "umi.bin")
    crec = orc.umi_cluster_batch(u["matrix"], u["job_offsets"], u["out_offsets"], 2)
    files.append(write_driver_file(tmp_path / "cluster.bin", 6, [int(u["umi_len"]), len(u["job_offsets"]) - 1, len(u["umis"]), len(u["matrix"]), 2,
                                                               u["umis"], u["job_offsets"], u["out_offsets"], u["matrix"], crec]) or tmp_path / "cluster.bin")
    au, aoff = pkg_synth_assign_batch()
    am, aoo = orc.umi_matrix_batch(au, aoff)
    aqv = (np.arange(len(aoff) - 1) % 2).astype(np.uint8)
    arec = orc.umi_assign_batch(am, aoff, aoo, None, aqv)
    files.append(write_driver_file(tmp_path / "assign.bin", 7, [12, len(aoff) - 1, len(au), au, aoff, aqv, arec]) or tmp_path / "assign.bin")
    for name in ("guided_umi_ed2", "guided_bc_mixed"):
        z = np.load(os.path.join(gold, name + ".npz"))
        raw = z["raw"].view(orc.GUIDED_HIT).reshape(len(z["slices"]), -1)
        files.append(write_driver_file(tmp_path / (name + ".bin"), 5, [
            int(z["L"]), int(z["bc"]), int(z["pm"]), int(z["post_len"]), int(z["bailout"]), int(z["slice_len"]), raw.shape[1],
            len(z["group_offsets"]) - 1, len(z["group_keys"]), len(z["all_keys"]), len(z["empty_keys"]), len(z["slices"]), z["group_keys"],
            z["group_offsets"], z["all_keys"], z["empty_keys"], z["slices"], z["anchor"], z["group_id"], z["ed"].astype(np.int32), z["result"],
            raw]) or tmp_path / (name + ".bin"))
    return files


def test_c_driver_fails_loudly_without_gpu(pkg, orc, tmp_path):
    import subprocess
    import torch
    exe = build_c(str(tmp_path / "abi_driver"), os.path.join(ROOT, "tests", "c_driver", "abi_driver.c"), [])
    files = driver_inputs(tmp_path, orc)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, str(files[0])], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_replays_golden_buffers(pkg, orc, tmp_path):
    """a C program (no Python, no torch in the process) gets bit-identical records through the C ABI"""
    import subprocess
    exe = build_c(str(tmp_path / "abi_driver"), os.path.join(ROOT, "tests", "c_driver", "abi_driver.c"), [])
    for f in driver_inputs(tmp_path, orc):
        r = subprocess.run([exe, str(f)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (f, r.stdout, r.stderr)


def host_driver_inputs(tmp_path):
    """binary dumps of the interpreter-run vectors of the host-side entry points (tests/c_driver/host_driver.c reads them)"""
    import numpy as np
    import __graft_entry__ as g
    import workloads
    pkg = g.load_package()
    gold = os.path.join(ROOT, "tests", "golden")
    i32, i64 = (lambda v: np.int32(v).tobytes()), (lambda v: np.int64(v).tobytes())
    z = np.load(os.path.join(gold, "ref_grouper.npz"))
    off = z["offsets"]
    with open(tmp_path / "grouper.bin", "wb") as f:
        f.write(b"SLRH" + np.uint32(1).tobytes() + i64(len(off) - 1))
        for c in range(len(off) - 1):
            a, b = int(off[c]), int(off[c + 1])
            f.write(i64(b - a) + i32(z["max_dist"][c]) + i64(z["id_before"][c]) + i64(z["id_after"][c]) + i32(z["keep_data_end"][c]) +
                    i32(bool(str(z["thrown"][c]))) + i64(z["n_done"][c]))
            for k, dt in (("position", np.int32), ("flags", np.int32), ("has_position", np.uint8), ("region_in", np.int64), ("region_out", np.int64)):
                f.write(np.ascontiguousarray(z[k][a:b], dtype=dt).tobytes())
    z = np.load(os.path.join(gold, "ref_jobs.npz"))
    off, joff = z["offsets"], z["job_offsets"]
    with open(tmp_path / "jobs.bin", "wb") as f:
        f.write(b"SLRH" + np.uint32(2).tobytes() + i64(len(off) - 1))
        for c in range(len(off) - 1):
            bc, region = z["barcode"][off[c]:off[c + 1]], z["region"][off[c]:off[c + 1]]
            valid = ((bc >= 0) & (region >= 0)).astype(np.uint8)
            # the recorded set of jobs, laid out in the library's documented order: groups by ascending (barcode, region), parts of a split group in sequence
            jobs = [z["job_reads"][joff[j]:joff[j + 1]] for j in np.nonzero(z["job_case"] == c)[0]]
            jobs.sort(key=lambda ids: (int(np.uint64(bc[ids[0]])), int(region[ids[0]]), int(ids[0])))
            o = np.concatenate([[0], np.cumsum([len(j) for j in jobs])]).astype(np.int64)
            f.write(i64(len(bc)) + i64(z["ram"][c]) + i64(len(jobs)) + bc.astype(np.uint64).tobytes() + region.astype(np.int64).tobytes() + valid.tobytes() +
                    o.tobytes() + (np.concatenate(jobs).astype(np.int64).tobytes() if jobs else b""))
    z = np.load(os.path.join(gold, "ref_needleman.npz"))
    with open(tmp_path / "needleman.bin", "wb") as f:
        f.write(b"SLRH" + np.uint32(3).tobytes() + i64(len(z["L"])))
        for i in range(len(z["L"])):
            f.write(np.uint64(workloads.g_pack(str(z["template"][i]))).tobytes() + np.uint64(workloads.g_pack(str(z["read"][i]))).tobytes() +
                    i32(z["L"][i]) + i32(z["custom"][i]) + z["scores"][i].astype(np.int32).tobytes() + z["counts"][i].astype(np.int32).tobytes())
    z = np.load(os.path.join(gold, "ref_usedlist.npz"))
    off = z["offsets"]
    with open(tmp_path / "usedlist.bin", "wb") as f:
        f.write(b"SLRH" + np.uint32(4).tobytes() + i64(len(off) - 1))
        for c in range(len(off) - 1):
            a, b = int(off[c]), int(off[c + 1])
            f.write(i64(b - a) + i32(z["ed"][c]) + i32(z["min_count_fold"][c]) + i32(z["cells_fold"][c]) + i64(z["record_count"][c]) +
                    z["barcodes"][a:b].astype(np.uint64).tobytes() + z["counts"][a:b].astype(np.int32).tobytes() + np.ascontiguousarray(z["collide"][a:b]).tobytes() +
                    z["kept"][a:b].astype(np.uint8).tobytes() + z["count_filter_keep"][a:b].astype(np.uint8).tobytes())
    return [tmp_path / "grouper.bin", tmp_path / "jobs.bin", tmp_path / "needleman.bin", tmp_path / "usedlist.bin"]


def test_c_host_driver_replays_reference_vectors(pkg, tmp_path):
    """a C program without JVM, Python or GPU forms the reference's regions, jobs, Needleman counts and used-barcode lists through the C ABI's host-side entry points"""
    import subprocess
    exe = build_c(str(tmp_path / "host_driver"), os.path.join(ROOT, "tests", "c_driver", "host_driver.c"), [])
    for f in host_driver_inputs(tmp_path):
        r = subprocess.run([exe, str(f)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (f, r.stdout, r.stderr)
