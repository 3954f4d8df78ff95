"""The ncu-derived per-kernel counts bench.py's roofline uses (profiles/r*_kernel_profile.json) describe the sources in the tree: every entry that
carries a src_sha256 must equal the hash of its kernel's translation unit + transitive includes (csrc/ and include/sicelore_gpu.h).  A kernel
or record-layout edit without a new profile would silently drop the roofline from the bench line — this fails here, on the CPU, instead."""
import glob
import json
import os

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_latest_kernel_profile_matches_the_sources():
    pkg = g.load_package()
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernel_profile.json")))
    assert files
    prof = json.load(open(files[-1]))
    checked = 0
    for k in prof["kernels"]:
        want = pkg.kernel_src_sha256(k["kernel"])
        if k.get("src_sha256") and want:
            assert k["src_sha256"] == want, "%s: %s was edited after %s was captured" % (k["kernel"], pkg.KERNEL_TU, os.path.basename(files[-1]))
            checked += 1
    assert checked >= 4


def test_bench_finds_the_profile_of_the_dominant_kernel():
    import bench
    pkg = g.load_package()
    for kernel in ("bc_assign_kernel<2>", "bc_assign_kernel<1>"):
        p = bench.load_profile(pkg, kernel)
        assert p and p["inst_executed_per_unit"] > 0 and p["file"].startswith("profiles/")
