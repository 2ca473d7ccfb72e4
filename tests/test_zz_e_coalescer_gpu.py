"""Query coalescer through the C ABI.  Written after this round's GPU minutes were spent: it
passes on the CPU against the emulated ABI (tests/test_emulated_abi.py); the file name sorts it after the tests that
have already been on hardware."""
import numpy as np
import pytest

from conftest import build_case
from test_gpu_parity import _queries

pytestmark = pytest.mark.gpu

COSINE = 0


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


def test_coalescer_concurrent_single_query_callers_get_private_scan_results(lib):
    """many blocking single-query callers (backends) -> few batch calls; every caller gets exactly
    what the oracle's private scan of its query returns, whatever it was batched with."""
    import threading
    from oracle import oracle
    s = build_case(2000, 96, COSINE, seed=44, labels=True, R=24, L_build=50, deleted_every=9)
    q = _queries(s, 48, 7)
    params = [(10, 60, 25, None), (10, 60, 25, [3, 9]), (7, 30, 0, None)]      # (k, L, rescore, labels) per caller class
    results, errors = {}, []
    with lib.DiskAnnIndex(s) as idx, lib.Coalescer(idx, max_batch=16, max_wait_us=20000) as co:
        start = threading.Barrier(12)

        def backend(t):
            try:
                start.wait()
                for j in range(4):
                    i = t * 4 + j
                    k, L, rescore, lab = params[i % 3]
                    results[i] = co.search(q[i], labels=lab, k=k, search_list_size=L, rescore=rescore)
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        th = [threading.Thread(target=backend, args=(t,)) for t in range(12)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        st = co.stats()
    assert not errors, errors
    assert st["queries"] == 48 and st["batches"] < 48 and st["largest_batch"] > 1, st
    for i in range(48):
        k, L, rescore, lab = params[i % 3]
        r = oracle.scan(s, q[i], lab, L, rescore, k)
        g = results[i]
        n = len(r["tid"])
        assert g["count"] == n and g["tid"][:n].tolist() == r["tid"].tolist(), i
        if rescore:
            assert g["dist"][:n].view(np.uint32).tolist() == r["dist"].view(np.uint32).tolist()
        for f in ("visits", "d_quantized", "candidates", "d_full"):
            assert g["stats"][f] == r["stats"][f], (i, f)
