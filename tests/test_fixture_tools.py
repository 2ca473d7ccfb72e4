"""Host-side pieces of the benchmark fixture (tools/fixture.py) that need no GPU: the lazily committed row table the
CPU oracle reads its rerank rows from, the tid <-> node bijection, the usable-core count, and the rule that tells which
heap rows the reference algorithm fetches (checked against the oracle's own stream)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import build_case  # noqa: E402
from oracle import fixtures, oracle  # noqa: E402
from tools import fixture as fx  # noqa: E402


def test_sparse_rows_commit_only_what_is_written():
    n, dim = 5_000_000, 768            # 15 GB of address space, a few MB touched
    rows = fx.SparseRows(n, dim)
    assert rows.arr.shape == (n, dim) and rows.arr.dtype == np.float32
    idx = np.array([0, 17, 4_999_999, 123_456, 17], np.int64)
    miss = rows.missing(idx)
    assert miss.tolist() == [0, 17, 123_456, 4_999_999]
    vals = np.arange(len(miss) * dim, dtype=np.float32).reshape(len(miss), dim)
    rows.put(miss, vals)
    assert rows.missing(idx).size == 0
    assert np.array_equal(rows.arr[123_456], vals[2]) and np.array_equal(rows.arr[4_999_999], vals[3])
    assert not rows.arr[1].any()       # a row nobody supplied reads as zeros


def test_tid_node_bijection_and_invalid_rows():
    from pgvectorscale_b200.snapshot import make_heap_tids
    t = make_heap_tids(1001)
    assert np.array_equal(fx.tid_to_node(t), np.arange(1001))
    assert fx.tid_to_node(np.array([0xFFFFFFFFFFFFFFFF], np.uint64)).tolist() == [-1]


def test_host_cores_respects_affinity_and_quota():
    c = fx.host_cores()
    assert 1 <= c["effective"] <= c["affinity"] <= c["logical"]
    if c["cgroup_quota"] is not None:
        assert c["effective"] <= int(np.ceil(c["cgroup_quota"]))


@pytest.mark.parametrize("rescore", [0, 7, 40])
def test_rerank_rows_are_exactly_the_rows_the_oracle_reads(rescore):
    """Give the oracle a row table that holds ONLY the rows oracle_rerank_rows names (all others zero): its results
    must equal a scan over the full table - i.e. the rule 'first rescore + k - 1 stream items' covers every heap fetch."""
    s = build_case(1500, 48, 0, seed=5, R=16, L_build=32, deleted_every=11)
    q = fixtures.gen_vectors(12, 48, 9, "normal")
    k, L = 5, 30
    full = oracle.scan_batch(s, q, None, None, L, rescore, k, threads=1)
    need = fx.oracle_rerank_rows(oracle, s, q, L, rescore, k, threads=1)
    rows = fx.SparseRows(s.n, s.dim)
    rows.put(need, np.asarray(s.vectors)[need])
    keep = s.vectors
    try:
        s.vectors = rows.arr
        sparse = oracle.scan_batch(s, q, None, None, L, rescore, k, threads=1)
    finally:
        s.vectors = keep
    assert np.array_equal(full[0], sparse[0])
    assert np.array_equal(full[1].view(np.uint32), sparse[1].view(np.uint32))
    if rescore:
        assert 0 < len(need) <= len(q) * (rescore + k - 1)
    else:
        assert len(need) == 0
