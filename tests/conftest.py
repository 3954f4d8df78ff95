import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: pytest-timeout's marker (registered here too, so that a box without the plugin only ignores it)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    p = g.load_package()
    p.build()
    return p


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.build()
    return o


@pytest.fixture(scope="session")
def sim():
    """CPU replay of the kernels' per-lane code (tests/host_sim)."""
    import ctypes as C
    import __graft_entry__ as g
    L = C.CDLL(g.build_host_sim())
    L.sim_bc_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_longlong), C.c_void_p]
    L.sim_bc_collide.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_longlong)]
    L.sim_umi_dist.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    L.sim_guided_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    L.sim_guided_filter_skips.restype = C.c_longlong
    L.sim_guided_far_nodes.restype = C.c_longlong
    L.sim_guided_filter_violations.restype = C.c_longlong
    L.sim_guided_filter_violations.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_ulonglong]
    return L


@pytest.fixture(scope="session")
def ctx(pkg):
    return pkg.Context(0, n_streams=2)
