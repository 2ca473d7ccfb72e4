"""SURVEY.md §8d config 1 at its stated size (100k x 768-d uniform, serial reference build, one query, L = 100,
rescore = 50, k = 10; cosine / 2-bit and L2 / 1-bit): the CUDA path returns the committed golden rows - ids, rerank
distance bits, counters - through the batch call and through the scan operator, and the oracle rebuilt here still
gives the committed answer (stream ids and a checksum of the neighbour lists included)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _build(name):
    from golden.make_config1_golden import answer, make_case
    s, q = make_case(name)
    return name, s, q, answer(s, q)


@pytest.fixture(scope="module")
def cases():
    from golden.make_config1_golden import CASES
    with mp.get_context("fork").Pool(len(CASES)) as pool:      # the serial builds side by side: ~100 s instead of ~200 s
        return {name: (s, q, a) for name, s, q, a in pool.map(_build, list(CASES))}


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


@pytest.mark.parametrize("name", ["cosine_b2", "l2_b1"])
def test_config1_at_size_matches_golden(lib, cases, name):
    from golden.make_config1_golden import K, L, RESCORE
    z = np.load(os.path.join(HERE, "golden", "config1_golden.npz"))
    s, q, a = cases[name]
    for f in ("tid", "dist_bits", "stream", "counters", "nbrs_crc"):       # the oracle still gives the committed answer
        assert np.array_equal(a[f], z[f"{name}/{f}"]), f
    with lib.DiskAnnIndex(s) as idx:
        g = idx.search_batch(q, k=K, search_list_size=L, rescore=RESCORE)
        assert g["tid"][0].tolist() == z[f"{name}/tid"].tolist()
        assert g["dist"][0].view(np.uint32).tolist() == z[f"{name}/dist_bits"].tolist()
        visits, dq, cand, dfull, slen = (int(x) for x in z[f"{name}/counters"])
        st = g["stats"]
        assert (int(st["visits"][0]), int(st["d_quantized"][0]), int(st["candidates"][0]), int(st["d_full"][0])) == \
            (visits, dq, cand, dfull)
        sc = idx.begin_scan()
        sc.rescan(q[0], search_list_size=L, rescore=RESCORE)
        rows = [sc.gettuple() for _ in range(K)]
        assert [(r[0] << 16) | r[1] for r in rows] == z[f"{name}/tid"].tolist()
        assert np.array([r[3] for r in rows], np.float32).view(np.uint32).tolist() == z[f"{name}/dist_bits"].tolist()
        sc.end()
