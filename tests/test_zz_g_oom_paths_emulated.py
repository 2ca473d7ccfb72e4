"""Error paths of the host code under the emulated ABI only (the fake CUDA runtime can make the k-th cudaMalloc fail):
every allocation site of load / batch search / streaming scan / build must turn an out-of-memory into DANN_ERR_OOM (or
succeed), never crash, and leave the library usable.  Marked gpu so that it runs in the emulated-ABI subprocess
(tests/test_emulated_abi.py); on the real GPU box it skips itself - there is nothing to inject there."""
import ctypes as C

import numpy as np
import pytest

from conftest import build_case, emulating

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not emulating(), reason="needs the fake CUDA runtime's failure injection")]


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    return diskann


def _sequence(lib, s, q):
    """One of everything that allocates."""
    with lib.DiskAnnIndex(s) as idx:
        idx.search_batch(q, k=5, search_list_size=20, rescore=10)
        idx.search_batch(q, labels=[[1], [2, 3]], k=5, search_list_size=20, rescore=10)
        sc = idx.begin_scan()
        sc.rescan(q[0], search_list_size=15, rescore=5)
        for _ in range(8):
            sc.gettuple()
        sc.end()


def test_every_allocation_site_reports_oom_cleanly(lib):
    s = build_case(300, 32, 0, seed=3, kind="normal", R=12, L_build=24, labels=True)
    from oracle import fixtures
    q = fixtures.gen_vectors(2, 32, 5, "normal")
    inject = lib.load_library().fake_cuda_fail_malloc_after
    inject.argtypes = [C.c_long]
    inject.restype = None
    failures = 0
    for k in range(0, 400):
        inject(k)
        try:
            _sequence(lib, s, q)
            inject(-1)
            break                       # k is past the last allocation of the sequence: everything succeeded
        except lib.DiskAnnError as e:
            assert e.code == -4, (k, e)   # DANN_ERR_OOM, with a message naming the call
            failures += 1
        finally:
            inject(-1)
        _sequence(lib, s, q)            # and the library still works afterwards
    else:
        pytest.fail("the sequence never ran out of allocation sites")
    assert failures >= 20, failures


def test_a_cuda_error_poisons_the_handle_and_only_the_handle(lib):
    """A non-OOM CUDA error is sticky for that index (DANN_ERR_CUDA on every later call, include/diskann_b200.h), while a
    freshly loaded index works."""
    from oracle import fixtures
    s = build_case(200, 16, 1, seed=4, kind="normal", R=8, L_build=16)
    q = fixtures.gen_vectors(2, 16, 5, "normal")
    inject = lib.load_library().fake_cuda_fail_sync_after
    inject.argtypes = [C.c_long]
    inject.restype = None
    idx = lib.DiskAnnIndex(s)
    try:
        idx.search_batch(q, k=3, search_list_size=10, rescore=5)
        inject(0)
        with pytest.raises(lib.DiskAnnError) as e:
            idx.search_batch(q, k=3, search_list_size=10, rescore=5)
        assert e.value.code == -2
        inject(-1)
        with pytest.raises(lib.DiskAnnError, match="poisoned") as e:
            idx.search_batch(q, k=3, search_list_size=10, rescore=5)
        assert e.value.code == -2
    finally:
        inject(-1)
        idx.close()
    with lib.DiskAnnIndex(s) as fresh:
        assert fresh.search_batch(q, k=3, search_list_size=10, rescore=5)["count"].tolist() == [3, 3]


def test_builder_allocation_sites_report_oom_cleanly(lib):
    from pgvectorscale_b200.snapshot import INVALID_NODE
    s = build_case(120, 16, 1, seed=6, kind="normal", R=8, L_build=16)
    s.R = 64
    s.nbrs = np.full((120, 64), INVALID_NODE, np.uint32)
    s.start_default = 0
    inject = lib.load_library().fake_cuda_fail_malloc_after
    inject.argtypes = [C.c_long]
    inject.restype = None
    failures = 0
    with lib.DiskAnnIndex(s) as idx:
        for k in range(0, 200):
            inject(k)
            try:
                idx.build_graph(12, 24, 1.2, 64)
                inject(-1)
                break
            except lib.DiskAnnError as e:
                assert e.code == -4, (k, e)
                failures += 1
            finally:
                inject(-1)
        else:
            pytest.fail("the build never ran out of allocation sites")
        st = idx.build_graph(12, 24, 1.2, 64)          # and a clean build still works on the same handle
        nb = idx.download_nbrs()
    assert failures >= 10 and st["batches"] >= 3
    assert ((nb != 0xFFFFFFFF).sum(1)[1:] >= 1).all()
