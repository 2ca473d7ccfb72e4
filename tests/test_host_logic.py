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


def test_plain_storage_snapshot_has_no_cpu_route_either():
    """The plain layout is served by the CUDA path only: without a device the load fails with
    DANN_ERR_NO_DEVICE like any other snapshot - it is never routed to the oracle."""
    import pytest
    from conftest import build_case
    from oracle import fixtures
    from pgvectorscale_b200 import diskann
    if diskann.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    s = fixtures.to_plain(build_case(64, 16, 1, seed=2, R=8, L_build=16))
    with pytest.raises(diskann.DiskAnnError) as e:
        diskann.DiskAnnIndex(s)
    assert e.value.code == -3


def test_raw_snapshot_file_round_trips_through_the_c_reader(tmp_path):
    """Snapshot.save_raw -> harness/snapshot_raw.h (C99): every header field and array arrives intact."""
    import os
    import shutil
    import subprocess
    from conftest import build_case
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("gcc not available")
    s = build_case(200, 24, 1, seed=3, R=8, L_build=16, labels=True, deleted_every=7)
    s.save_raw(str(tmp_path / "s.raw"))
    (tmp_path / "t.c").write_text(r'''
#include "snapshot_raw.h"
int main(int argc, char **argv) {
    dann_snapshot_desc s; const float *iv = 0;
    (void)argc;
    void *b = dann_snapshot_raw_read(argv[1], &s, &iv);
    if (!b) return 1;
    unsigned long long h = 1469598103934665603ull;
    #define MIX(p, bytes) for (size_t i = 0; i < (size_t)(bytes); i++) h = (h ^ ((const unsigned char *)(p))[i]) * 1099511628211ull
    MIX(s.mean, s.dim_index * 4); MIX(s.m2, s.dim_index * 4); MIX(s.codes, (size_t)s.n * s.words * 8);
    MIX(s.nbrs, (size_t)s.n * s.R * 4); MIX(s.heap_tid, (size_t)s.n * 8); MIX(s.vectors, (size_t)s.n * s.dim * 4);
    MIX(s.start_labels, s.n_start_labels * 2); MIX(s.start_label_nodes, s.n_start_labels * 4);
    MIX(s.label_off, ((size_t)s.n + 1) * 4); MIX(s.labels, (size_t)s.label_off[s.n] * 2);
    printf("%u %u %u %u %u %u %d %d %llu %u %u %d %llu\n", s.n, s.dim, s.dim_index, s.bits, s.words, s.R, s.distance_type,
           s.has_labels, (unsigned long long)s.count, s.start_default, s.n_start_labels, iv != 0, h);
    free(b); return 0;
}''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "t")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                    "-I" + os.path.join(root, "harness"), str(tmp_path / "t.c"), "-o", exe], check=True)
    out = subprocess.run([exe, str(tmp_path / "s.raw")], capture_output=True, text=True, check=True).stdout.split()
    h = 1469598103934665603
    for a, dt in ((s.mean, np.float32), (s.m2, np.float32), (s.codes, np.uint64), (s.nbrs, np.uint32), (s.heap_tid, np.uint64),
                  (s.vectors, np.float32), (s.start_labels, np.int16), (s.start_label_nodes, np.uint32),
                  (s.label_off, np.uint32), (s.labels, np.int16)):
        for byte in np.ascontiguousarray(a, dtype=dt).tobytes():
            h = ((h ^ byte) * 1099511628211) % (1 << 64)
    want = [s.n, s.dim, s.dim_index, s.bits, s.words, s.R, int(s.distance_type), 1, int(s.count), int(s.start_default),
            len(s.start_labels), 0, h]
    assert [int(x) for x in out] == want
    # a truncated file, an array whose length disagrees with the header geometry, and a length that would wrap the
    # bounds check are all refused (exit status 1 = NULL), never read past the buffer
    raw = (tmp_path / "s.raw").read_bytes()
    hdr_end = 8 + 16 * 8
    bad = {}
    bad["truncated"] = raw[: len(raw) // 2]
    n_up = bytearray(raw)
    n_up[8:16] = np.uint64(s.n + 1).tobytes()              # header says one more node than the arrays hold
    bad["geometry"] = bytes(n_up)
    wrap = bytearray(raw)
    wrap[hdr_end:hdr_end + 8] = np.uint64(2 ** 64 - 8).tobytes()   # off + len wraps around
    bad["wrap"] = bytes(wrap)
    plain_flag = bytearray(raw)
    plain_flag[8 + 11 * 8:8 + 12 * 8] = np.uint64(1).tobytes()     # storage_type = plain without index_vectors
    bad["storage_type"] = bytes(plain_flag)
    for name, blob in bad.items():
        (tmp_path / "bad.raw").write_bytes(blob)
        r = subprocess.run([exe, str(tmp_path / "bad.raw")], capture_output=True, text=True)
        assert r.returncode == 1, (name, r.returncode, r.stdout, r.stderr)
