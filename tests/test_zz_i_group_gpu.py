"""dann_group: one replica per GPU inside one process, the batch cut into contiguous slices, rows
back in query order.  The gathered result must equal the oracle - and the single-device call - row for row, also when
the batch does not divide evenly and under label keys.  Needs two GPUs on hardware (skipped otherwise); under the
emulated ABI (tests/test_emulated_abi.py, SIMT_FAKE_DEVICES=2) the two replicas are two fake devices."""
import numpy as np
import pytest

from conftest import build_case
from oracle import fixtures, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


def _need_two(lib):
    if lib.device_count() < 2:
        pytest.skip("needs two CUDA devices (or SIMT_FAKE_DEVICES=2 under the emulated ABI)")


def test_group_rows_equal_oracle_and_single_device(lib):
    _need_two(lib)
    s = build_case(3000, 96, 0, seed=7, R=24, L_build=48, deleted_every=13)
    q = fixtures.gen_vectors(37, 96, 11, "normal")          # 37 = 19 + 18: uneven slices
    with lib.IndexGroup(s, ndev=2) as g:
        assert g.size == 2
        got = g.search_batch(q, k=10, search_list_size=60, rescore=40)
        one = g.search_batch(q[:1], k=10, search_list_size=60, rescore=40)     # fewer queries than devices
    otid, odist, ocount, ostats = oracle.scan_batch(s, q, None, None, 60, 40, 10)
    assert np.array_equal(got["tid"], otid)
    assert np.array_equal(got["dist"].view(np.uint32), odist.view(np.uint32))
    assert np.array_equal(got["count"], ocount)
    assert np.array_equal(got["stats"]["visits"].astype(np.uint64), ostats["visits"])
    assert np.array_equal(one["tid"], otid[:1])
    with lib.DiskAnnIndex(s, device=0) as idx:
        single = idx.search_batch(q, k=10, search_list_size=60, rescore=40)
    assert np.array_equal(single["tid"], got["tid"])


def test_group_label_keys_follow_their_slices(lib):
    _need_two(lib)
    s = build_case(2500, 64, 1, seed=21, labels=True, R=24, L_build=48)
    q = fixtures.gen_vectors(21, 64, 5, "normal")
    rng = np.random.default_rng(3)
    labs = [[int(x) for x in rng.integers(1, 17, size=1 + (i % 2))] for i in range(21)]
    with lib.IndexGroup(s, devices=[1, 0]) as g:                # replica order is the caller's
        got = g.search_batch(q, labels=labs, k=10, search_list_size=50, rescore=30)
    flat = np.array([x for ls in labs for x in ls], np.int16)
    off = np.cumsum([0] + [len(ls) for ls in labs]).astype(np.int32)
    otid, odist, _, _ = oracle.scan_batch(s, q, flat, off, 50, 30, 10)
    assert np.array_equal(got["tid"], otid)
    assert np.array_equal(got["dist"].view(np.uint32), odist.view(np.uint32))


def test_group_create_rejects_a_missing_device(lib):
    _need_two(lib)
    s = build_case(200, 32, 1, seed=3, R=8, L_build=16)
    with pytest.raises(lib.DiskAnnError):
        lib.IndexGroup(s, devices=[0, 99])
