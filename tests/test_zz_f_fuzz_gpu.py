"""Randomised differential test through the C ABI: small random indexes and scan parameters at the edges (1-node
graphs, lists of 1 id or more than 64, 1-dimensional vectors, every tuple deleted, k > n, label keys with no start
node ...) against the oracle - batch call and streaming scan operator.  Cheap on the B200 and on the CPU (it runs
against the emulated ABI in tests/test_emulated_abi.py).  Written after this round's GPU minutes were spent; the file
name sorts it after the tests that have been on hardware."""
import numpy as np
import pytest

from conftest import build_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


def _case(rng):
    n = int(rng.choice([1, 2, 3, 5, 17, 64, 65, 300]))
    dim = int(rng.choice([1, 2, 3, 8, 31, 33, 64, 100]))
    dist = int(rng.integers(0, 3))
    bits = int(rng.choice([1, 2]))
    R = int(rng.choice([1, 2, 7, 16, 33, 64, 70]))
    dim_index = None if dim < 3 or rng.random() < 0.6 else int(rng.integers(1, dim))
    labels = bool(rng.random() < 0.4)
    deleted_every = int(rng.choice([0, 0, 3, 2, 1]))
    return dict(n=n, dim=dim, dist=dist, bits=bits, R=R, dim_index=dim_index, labels=labels, deleted_every=deleted_every)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DANN_FUZZ_SEEDS", "40"))))
def test_random_small_index_batch_and_scan_equal_oracle(lib, seed):
    from oracle import fixtures, oracle
    rng = np.random.default_rng(9000 + seed)
    c = _case(rng)
    s = build_case(c["n"], c["dim"], c["dist"], bits=c["bits"], seed=seed, kind=str(rng.choice(["normal", "uniform"])),
                   R=c["R"], L_build=max(2 * c["R"], 8), labels=c["labels"], dim_index=c["dim_index"],
                   deleted_every=c["deleted_every"])
    B = int(rng.integers(1, 6))
    q = fixtures.gen_vectors(B, c["dim"], 100 + seed, "normal")
    if rng.random() < 0.3:
        q[0] = 0.0                                         # zero vector: the cosine normalisation's epsilon branch
    keys = None
    if c["labels"] and rng.random() < 0.7:
        keys = [[int(x) for x in rng.integers(1, 20, size=int(rng.integers(0, 4)))] for _ in range(B)]
    with lib.DiskAnnIndex(s) as idx:
        for _ in range(2):
            k = int(rng.choice([1, 3, 20]))
            L = int(rng.choice([1, 2, 5, 50]))
            rescore = int(rng.choice([0, 1, 7, 100]))
            g = idx.search_batch(q, labels=keys, k=k, search_list_size=L, rescore=rescore)
            for b in range(B):
                r = oracle.scan(s, q[b], None if keys is None else keys[b], L, rescore, k)
                nrow = len(r["tid"])
                assert int(g["count"][b]) == nrow, (c, k, L, rescore, b)
                assert g["tid"][b, :nrow].tolist() == r["tid"].tolist(), (c, k, L, rescore, b)
                if rescore:
                    assert g["dist"][b, :nrow].view(np.uint32).tolist() == r["dist"].view(np.uint32).tolist()
                for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len"):
                    assert int(g["stats"][f][b]) == r["stats"][f], (c, k, L, rescore, b, f)
        # the streaming operator on one query, to the end of the scan
        L, rescore = int(rng.choice([1, 4, 30])), int(rng.choice([0, 2, 50]))
        key = None if keys is None else keys[0]
        want = oracle.scan(s, q[0], key, L, rescore, 10_000)
        sc = idx.begin_scan()
        sc.rescan(q[0], labels=key, search_list_size=L, rescore=rescore)
        got = []
        while True:
            row = sc.gettuple()
            if row is None:
                break
            got.append((row[0] << 16) | row[1])
            assert len(got) <= c["n"]
        assert got == want["tid"].tolist(), (c, L, rescore)
        st = sc.stats()
        for f in ("visits", "d_quantized", "candidates", "d_full"):
            assert st[f] == want["stats"][f], (c, L, rescore, f)
        sc.end()


@pytest.mark.parametrize("key,want", [([1], 2), ([], 0), ([3], 1), (None, 4)])
def test_reference_null_and_empty_labels_kat_through_the_abi(lib, key, want):
    """labels/filtering_tests.rs:23-110 (restated in tests/test_oracle_kats.py) through the C ABI: batch call and scan."""
    from test_oracle_kats import _null_and_empty_labels_index
    s = _null_and_empty_labels_index()
    q = np.zeros((1, 3), np.float32)
    with lib.DiskAnnIndex(s) as idx:
        g = idx.search_batch(q, labels=None if key is None else [key], k=10, search_list_size=100, rescore=50)
        assert int(g["count"][0]) == want
        sc = idx.begin_scan()
        sc.rescan(q[0], labels=key, search_list_size=100, rescore=50)
        n = 0
        while sc.gettuple() is not None:
            n += 1
        sc.end()
        assert n == want


@pytest.mark.parametrize("seed", range(8))
def test_random_medium_index_long_scans_equal_oracle(lib, monkeypatch, seed):
    """Bigger random cases: up to 5000 nodes x 768 dimensions, lists of up to 64 ids, L up to 300, 300 rerank rows,
    k = 50, shared-memory heap tops from 64 entries to "as large as fits", optional forced workspace growth - the heap
    spills, pages straddle leaf levels, several query slots share a block."""
    from oracle import fixtures, oracle
    rng = np.random.default_rng(50000 + seed)
    n = int(rng.choice([700, 2000, 5000]))
    dim = int(rng.choice([16, 64, 200, 768]))
    dist = int(rng.integers(0, 3))
    R = int(rng.choice([16, 32, 50, 64]))
    labels = bool(rng.random() < 0.3)
    s = build_case(n, dim, dist, bits=int(rng.choice([1, 2])), seed=seed, kind="normal", R=R, L_build=2 * R, labels=labels,
                   deleted_every=int(rng.choice([0, 0, 7])))
    q = fixtures.gen_vectors(4, dim, 100 + seed, "normal")
    keys = None
    if labels and rng.random() < 0.6:
        keys = [[int(x) for x in rng.integers(1, 17, size=int(rng.integers(1, 3)))] for _ in range(4)]
    monkeypatch.setenv("DANN_SEARCH_HS", str(int(rng.choice([64, 512, 4096, 100000]))))
    if rng.random() < 0.5:
        monkeypatch.setenv("DANN_DEBUG_SHRINK", "4")
    k, L, rescore = int(rng.choice([10, 50])), int(rng.choice([50, 150, 300])), int(rng.choice([0, 50, 300]))
    with lib.DiskAnnIndex(s) as idx:
        g = idx.search_batch(q, labels=keys, k=k, search_list_size=L, rescore=rescore)
    for b in range(4):
        r = oracle.scan(s, q[b], None if keys is None else keys[b], L, rescore, k)
        nrow = len(r["tid"])
        assert int(g["count"][b]) == nrow and g["tid"][b, :nrow].tolist() == r["tid"].tolist(), (seed, b)
        if rescore:
            assert g["dist"][b, :nrow].view(np.uint32).tolist() == r["dist"].view(np.uint32).tolist()
        for f in ("visits", "d_quantized", "candidates", "d_full"):
            assert int(g["stats"][f][b]) == r["stats"][f], (seed, b, f)
