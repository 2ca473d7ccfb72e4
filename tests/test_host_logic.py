"""Host-side logic that needs no GPU: snapshot container, default parameters."""
import numpy as np

from conftest import build_case
from pgvectorscale_b200 import snapshot


def test_code_words_and_default_bits():
    # sbq/quantize.rs:38-46, meta_page.rs:312-323
    assert snapshot.code_words(768, 2) == 24 and snapshot.code_words(768, 1) == 12
    assert snapshot.code_words(65, 1) == 2 and snapshot.code_words(64, 1) == 1
    assert snapshot.default_bits(768) == 2 and snapshot.default_bits(899) == 2
    assert snapshot.default_bits(900) == 1 and snapshot.default_bits(1536) == 1


def test_snapshot_roundtrip(tmp_path):
    s = build_case(120, 40, 0, seed=2, R=8, L_build=16, labels=True)
    s.validate()
    p = str(tmp_path / "snap.npz")
    s.save(p)
    t = snapshot.Snapshot.load(p)
    t.validate()
    for f in ("n", "dim", "dim_index", "bits", "words", "R", "distance_type", "has_labels", "count",
              "start_default"):
        assert getattr(s, f) == getattr(t, f), f
    for f in ("mean", "m2", "codes", "nbrs", "heap_tid", "vectors", "start_labels",
              "start_label_nodes", "label_off", "labels"):
        assert np.array_equal(getattr(s, f), getattr(t, f)), f


def test_heap_tids_have_valid_offsets():
    t = snapshot.make_heap_tids(1000)
    assert ((t & np.uint64(0xFFFF)) >= 1).all() and len(set(t.tolist())) == 1000


def test_bench_keeps_stdout_for_the_json_line():
    """bench.py points fd 1 at stderr so that only its final JSON line reaches the real stdout
    (libraries such as NCCL print banners there)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stdout == ""
    assert "--impl" in r.stderr and "--gpus" in r.stderr


def test_shard_bounds_and_recall_helper_shapes():
    from pgvectorscale_b200.group import shard_bounds
    assert shard_bounds(1024, 8, 0) == (0, 128) and shard_bounds(1024, 8, 7) == (896, 1024)
    assert shard_bounds(10, 4, 1) == (3, 6)


def test_plain_storage_snapshot_is_refused_not_emulated():
    """The plain layout exists in the oracle only: the product path must refuse it rather
    than route anywhere else."""
    import pytest
    from conftest import build_case
    from oracle import fixtures
    from pgvectorscale_b200.diskann import DiskAnnError, DiskAnnIndex
    s = fixtures.to_plain(build_case(64, 16, 1, seed=2, R=8, L_build=16))
    with pytest.raises(DiskAnnError, match="plain"):
        DiskAnnIndex(s)
