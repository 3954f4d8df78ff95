"""The C-ABI library loads without a GPU and exports exactly what include/sicelore_gpu.h declares; compute calls fail
loudly (SLR_E_NODEVICE) instead of falling back to the CPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sicelore_gpu.h")).read()
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
