"""GPU parity: the CUDA scan path (through the C ABI) against the CPU oracle.

Bar: bit-exact row ids (TIDs), bit-exact rerank distances, equal counters
(visits / d_quantized / candidates / d_full / stream_len)."""
import numpy as np
import pytest

from conftest import buffer_device, build_case, dptr, emulating

pytestmark = pytest.mark.gpu

COSINE, L2, IP = 0, 1, 2


def _queries(s, B, seed, kind="normal"):
    from oracle import fixtures
    return fixtures.gen_vectors(B, s.dim, seed, kind)


def _compare_batch(s, idx, q, k, L, rescore, labels=None):
    from oracle import oracle
    B = q.shape[0]
    if labels is None:
        lab = off = None
    else:
        off = np.zeros(B + 1, np.int32)
        vals = []
        for i, ls in enumerate(labels):
            vals.extend(ls)
            off[i + 1] = len(vals)
        lab = np.asarray(vals, np.int16)
    otid, odist, ocount, ostats = oracle.scan_batch(s, q, lab, off, L, rescore, k, threads=0)
    g = idx.search_batch(q, labels=labels, k=k, search_list_size=L, rescore=rescore)
    assert np.array_equal(g["count"], ocount)
    assert np.array_equal(g["tid"], otid), "row ids differ from the oracle"
    if rescore > 0:
        # exact rerank distances must be the same f32 bit patterns
        assert np.array_equal(g["dist"].view(np.uint32), odist.view(np.uint32))
    for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len"):
        assert np.array_equal(g["stats"][f].astype(np.uint64), ostats[f]), f
    assert not g["stats"]["status"].any()
    return g


@pytest.fixture(scope="module")
def lib(lib_built):
    from pgvectorscale_b200 import diskann
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return diskann


@pytest.mark.parametrize("dist,bits,kind", [(COSINE, 2, "normal"), (COSINE, 1, "normal"),
                                            (L2, 2, "uniform"), (IP, 1, "uniform")])
def test_batch_768d_matches_oracle(lib, dist, bits, kind):
    s = build_case(3000, 768, dist, bits=bits, seed=11 + dist + bits, kind=kind)
    with lib.DiskAnnIndex(s) as idx:
        q = _queries(s, 64, 77, kind)
        _compare_batch(s, idx, q, k=10, L=100, rescore=50)
        _compare_batch(s, idx, q[:16], k=10, L=25, rescore=0)     # rescore=0 bypasses rerank
        _compare_batch(s, idx, q[:16], k=37, L=10, rescore=7)
        _compare_batch(s, idx, q[:8], k=100, L=200, rescore=200)


@pytest.mark.parametrize("dim,dist", [(100, L2), (37, COSINE), (130, IP), (1537, COSINE)])
def test_odd_dimensions_tail_path(lib, dim, dist):
    """dims that are not multiples of 32 / 4: scalar tail of the AVX2 body, unaligned rows."""
    s = build_case(800, dim, dist, seed=5 + dim, kind="uniform", R=20, L_build=40)
    with lib.DiskAnnIndex(s) as idx:
        q = _queries(s, 32, 99, "uniform")
        _compare_batch(s, idx, q, k=10, L=50, rescore=20)


def test_labeled_filter_matches_oracle(lib):
    s = build_case(2500, 128, COSINE, seed=21, labels=True, R=32, L_build=64)
    rng = np.random.default_rng(3)
    B = 48
    q = _queries(s, B, 5)
    with lib.DiskAnnIndex(s) as idx:
        one = [[int(rng.integers(1, 17))] for _ in range(B)]
        _compare_batch(s, idx, q, 10, 100, 50, labels=one)
        two = [[int(x) for x in rng.integers(1, 17, size=2)] for _ in range(B)]   # dup + unsorted
        _compare_batch(s, idx, q, 10, 60, 30, labels=two)
        # empty key => has_label_filter false but no start nodes => no rows; unknown label => no rows
        g = _compare_batch(s, idx, q[:4], 10, 100, 50, labels=[[], [999], [], [-5]])
        assert not g["count"].any()
        # unkeyed scan over a labeled index
        _compare_batch(s, idx, q[:16], 10, 100, 50)


def test_deleted_tuples_and_truncated_dims(lib):
    s = build_case(2000, 256, COSINE, seed=31, dim_index=128, deleted_every=3, R=32, L_build=64)
    with lib.DiskAnnIndex(s) as idx:
        q = _queries(s, 32, 6)
        g = _compare_batch(s, idx, q, 10, 100, 50)
        assert ((g["tid"] & np.uint64(0xFFFF)) != 0).all()


def test_small_and_degenerate_indexes(lib):
    from oracle import oracle
    # fewer nodes than k / rescore: the scan returns every live row then ends
    s = build_case(7, 64, L2, seed=41, kind="uniform", R=10, L_build=10)
    with lib.DiskAnnIndex(s) as idx:
        q = _queries(s, 5, 8, "uniform")
        g = _compare_batch(s, idx, q, 20, 100, 50)
        assert (g["count"] == 7).all()
    # index created on an empty table then filled: means = 0, count = 0 (build.rs:1419-1473)
    s = build_case(300, 32, COSINE, seed=43, kind="uniform", R=10, L_build=20, train_on_data=False)
    with lib.DiskAnnIndex(s) as idx:
        q = _queries(s, 8, 9, "uniform")
        _compare_batch(s, idx, q, 10, 100, 0)
        _compare_batch(s, idx, q, 10, 100, 50)
    # empty index
    s = build_case(0, 16, L2, seed=1)
    with lib.DiskAnnIndex(s) as idx:
        g = idx.search_batch(np.zeros((3, 16), np.float32), k=5)
        assert not g["count"].any() and (g["tid"] == lib.INVALID_TID).all()


def test_scan_operator_streams_like_amgettuple(lib):
    from oracle import oracle
    s = build_case(1500, 96, COSINE, seed=51, labels=True, R=24, L_build=50)
    q = _queries(s, 3, 12)
    with lib.DiskAnnIndex(s) as idx:
        sc = idx.begin_scan()
        for labels in (None, [3], [2, 9]):
            for qi in range(3):
                sc.rescan(q[qi], labels=labels, search_list_size=40, rescore=10)
                rows = []
                for _ in range(90):          # crosses the 16 -> 64 -> 256 refetch boundaries
                    r = sc.gettuple()
                    if r is None:
                        break
                    rows.append(r)
                ref = oracle.scan(s, q[qi], labels, 40, 10, 90)
                got_tid = np.array([(b << 16) | o for b, o, _, _ in rows], np.uint64)
                assert np.array_equal(got_tid, ref["tid"])
                assert np.array_equal(np.array([n for _, _, n, _ in rows], np.uint32), ref["node"])
                assert np.array_equal(np.array([d for *_, d in rows], np.float32).view(np.uint32),
                                      ref["dist"].view(np.uint32))
        # NULL order-by argument: zero vector, no labels (labels/mod.rs:214-216); all rows come back
        sc.rescan(None, labels=[1], search_list_size=100, rescore=50)
        n = 0
        while sc.gettuple() is not None:
            n += 1
        ref = oracle.scan(s, None, None, 100, 50, 5000)
        assert n == len(ref["tid"]) and n > s.n // 2
        sc.end()


def test_workspace_growth_retry_is_invisible(lib, monkeypatch):
    """A deliberately tiny first workspace forces the overflow -> rerun path."""
    s = build_case(3000, 768, COSINE, bits=2, seed=13)
    q = _queries(s, 16, 78)
    with lib.DiskAnnIndex(s) as idx:
        monkeypatch.setenv("DANN_SEARCH_HS", "256")      # heap mostly in the HBM tail
        _compare_batch(s, idx, q, 10, 100, 50)
        monkeypatch.delenv("DANN_SEARCH_HS")
        monkeypatch.setenv("DANN_DEBUG_SHRINK", "16")    # candidates / hash / visited all overflow
        _compare_batch(s, idx, q, 10, 100, 50)
        assert idx.last_batch_timing()["retries"] >= 1
        monkeypatch.delenv("DANN_DEBUG_SHRINK")


def test_single_warp_kernel_and_wide_lists(lib, monkeypatch):
    """Both search kernels give the oracle's answer: the two-warp kernel is the default for
    R <= 64; DANN_SEARCH_KERNEL=1 forces the single-warp kernel, which R > 64 always uses."""
    s = build_case(2500, 128, COSINE, seed=21, labels=True, R=32, L_build=64)
    q = _queries(s, 24, 5)
    with lib.DiskAnnIndex(s) as idx:
        monkeypatch.setenv("DANN_SEARCH_KERNEL", "1")
        _compare_batch(s, idx, q, 10, 100, 50)
        _compare_batch(s, idx, q, 10, 40, 20, labels=[[1 + (i % 16), 5] for i in range(24)])
        monkeypatch.setenv("DANN_SEARCH_BITMAP", "0")         # CAS hash-set flavour of the inserted-set
        _compare_batch(s, idx, q, 10, 100, 50)
        monkeypatch.delenv("DANN_SEARCH_KERNEL")
        _compare_batch(s, idx, q, 10, 100, 50)
        _compare_batch(s, idx, q, 10, 40, 20, labels=[[1 + (i % 16), 5] for i in range(24)])
        monkeypatch.delenv("DANN_SEARCH_BITMAP")
        # all three heap-entry layouts (4 B dist11|seq21, 4 B dist16|seq16, 8 B), both kernels
        for entry in ("0", "1", "2"):
            monkeypatch.setenv("DANN_SEARCH_ENTRY", entry)
            _compare_batch(s, idx, q[:12], 10, 100, 50)
            monkeypatch.setenv("DANN_SEARCH_KERNEL", "1")
            _compare_batch(s, idx, q[:12], 10, 60, 20)
            monkeypatch.delenv("DANN_SEARCH_KERNEL")
        monkeypatch.delenv("DANN_SEARCH_ENTRY")
    s = build_case(1200, 64, L2, seed=23, kind="uniform", R=70, L_build=80)
    with lib.DiskAnnIndex(s) as idx:
        _compare_batch(s, idx, _queries(s, 16, 6, "uniform"), 10, 60, 30)


def test_duplicate_ids_inside_a_neighbour_list(lib):
    """The reference's builder never repeats an id in a list, but the scan must not rely on it:
    the first occurrence (in list order) is the one that inserts."""
    s = build_case(1500, 96, COSINE, seed=29, R=40, L_build=60)
    rng = np.random.default_rng(4)
    nb = s.nbrs.copy()
    for i in range(0, s.n, 3):            # duplicate an early id late in the list (crosses the 32-lane chunks)
        deg = int((nb[i] != 0xFFFFFFFF).sum())
        if deg >= 36:
            nb[i, deg - 1] = nb[i, int(rng.integers(0, 8))]
            nb[i, 33] = nb[i, 34]
    s.nbrs = nb
    q = _queries(s, 24, 7)
    with lib.DiskAnnIndex(s) as idx:
        _compare_batch(s, idx, q, 10, 100, 50)


def test_sbq_distance_kernel(lib):
    import torch
    s = build_case(3000, 768, COSINE, bits=2, seed=13)
    with lib.DiskAnnIndex(s) as idx:
        dev = buffer_device()
        Q, npairs = 33, 200003
        rng = np.random.default_rng(0)
        qcodes = rng.integers(0, 2**63, size=(Q, idx.code_stride), dtype=np.uint64)
        qcodes[:, s.words:] = 0
        pq = rng.integers(0, Q, size=npairs, dtype=np.uint32)
        pn = rng.integers(0, s.n, size=npairs, dtype=np.uint32)
        d_q = torch.from_numpy(qcodes.view(np.int64)).to(dev)
        d_pq = torch.from_numpy(pq.view(np.int32)).to(dev)
        d_pn = torch.from_numpy(pn.view(np.int32)).to(dev)
        d_out = torch.empty(npairs, dtype=torch.int32, device=dev)
        idx.sbq_distance(dptr(d_q), dptr(d_pq), dptr(d_pn), dptr(d_out))
        x = s.codes[pn] ^ qcodes[pq][:, :s.words]
        ref = np.unpackbits(x.view(np.uint8), axis=1).sum(1).astype(np.int32)
        assert np.array_equal(d_out.cpu().numpy(), ref)


def test_prepare_and_full_distance_kernels(lib):
    import torch
    from oracle import oracle
    for dist, dim, dim_index in ((COSINE, 768, 768), (L2, 100, 64), (IP, 130, 130)):
        s = build_case(500, dim, dist, seed=61 + dim, kind="uniform", R=16, L_build=32, dim_index=dim_index)
        with lib.DiskAnnIndex(s) as idx:
            dev = buffer_device()
            B, m = 9, 21
            q = _queries(s, B, 14, "uniform") * 3.0
            q[0] = 0.0
            d_q = torch.from_numpy(q).to(dev)
            d_full = torch.empty((B, dim), dtype=torch.float32, device=dev)
            d_codes = torch.empty((B, idx.code_stride), dtype=torch.int64, device=dev)
            idx.prepare_queries(dptr(d_q), dptr(d_full), dptr(d_codes))
            full = d_full.cpu().numpy()
            codes = d_codes.cpu().numpy().view(np.uint64)
            for b in range(B):
                qf = oracle.preprocess_cosine(q[b]) if dist == COSINE else q[b]
                qi = q[b, :dim_index].copy()
                if dist == COSINE:
                    qi = oracle.preprocess_cosine(qi)
                assert np.array_equal(full[b].view(np.uint32), qf.view(np.uint32))
                ref = oracle.quantize(qi, s.bits, s.mean, s.m2, s.count)
                assert np.array_equal(codes[b, :s.words], ref)
                assert not codes[b, s.words:].any()
            rng = np.random.default_rng(1)
            nodes = rng.integers(0, s.n, size=(B, m), dtype=np.uint32)
            d_nodes = torch.from_numpy(nodes.view(np.int32)).to(dev)
            d_out = torch.empty((B, m), dtype=torch.float32, device=dev)
            idx.full_distance(dptr(d_full), dptr(d_nodes), dptr(d_out))
            out = d_out.cpu().numpy()
            for b in range(B):
                for i in range(m):
                    x = s.vectors[nodes[b, i]]
                    if dist == COSINE:
                        x = oracle.preprocess_cosine(x)
                    ref = np.float32(oracle.distance(dist, x, full[b], "avx2"))
                    assert out[b, i].view(np.uint32) == ref.view(np.uint32), (dist, b, i)


def test_argument_validation_does_not_poison_the_handle(lib):
    s = build_case(400, 64, L2, seed=71, kind="uniform", R=12, L_build=24)
    q = _queries(s, 4, 3, "uniform")
    with lib.DiskAnnIndex(s) as idx:
        for kw in (dict(k=0), dict(k=10, search_list_size=0), dict(k=10, search_list_size=10001),
                   dict(k=10, rescore=1001), dict(k=10, rescore=-1), dict(k=2_000_000, rescore=10)):
            with pytest.raises(lib.DiskAnnError) as e:
                idx.search_batch(q, **kw)
            assert e.value.code == -1, kw          # DANN_ERR_INVALID_ARG, not a CUDA error
        _compare_batch(s, idx, q, 10, 50, 20)      # the handle still works
        sc = idx.begin_scan()
        with pytest.raises(lib.DiskAnnError):
            sc.gettuple()                           # amgettuple before amrescan
        sc.end()


def test_index_without_vectors_then_set_vectors(lib):
    from oracle import oracle
    s = build_case(900, 96, COSINE, seed=73, R=16, L_build=32)
    vec = s.vectors
    s.vectors = None
    q = _queries(s, 8, 4)
    with lib.DiskAnnIndex(s) as idx:
        with pytest.raises(lib.DiskAnnError):
            idx.search_batch(q, k=5, rescore=10)    # no heap vectors yet: only rescore=0 scans
        s.vectors = vec
        g0 = idx.search_batch(q, k=5, rescore=0)
        otid, _, _, _ = oracle.scan_batch(s, q, None, None, 100, 0, 5)
        assert np.array_equal(g0["tid"], otid)
        idx.set_vectors(vec)
        _compare_batch(s, idx, q, 5, 100, 10)
        assert np.array_equal(idx.download_nbrs(), s.nbrs)


def test_scan_counters_after_every_gettuple(lib, monkeypatch):
    """The scan is a suspended search in HBM, resumed by each amgettuple: after row i its counters
    (visits / d_quantized / candidate / d_full, the ones amendscan logs, scan.rs:461-472) are what
    the reference's would be after i rows — not those of a larger prefetched LIMIT."""
    from oracle import oracle
    s = build_case(2000, 128, COSINE, seed=81, labels=True, R=24, L_build=50, deleted_every=9)
    q = _queries(s, 2, 15)
    with lib.DiskAnnIndex(s) as idx:
        sc = idx.begin_scan()
        for shrink in (None, "16"):                     # second pass: the workspace must regrow mid-scan
            if shrink:
                monkeypatch.setenv("DANN_DEBUG_SHRINK", shrink)
            for (labels, L, rescore) in ((None, 30, 12), ([4, 7], 20, 0), (None, 5, 3)):
                sc.rescan(q[0], labels=labels, search_list_size=L, rescore=rescore)
                for i in range(1, 41):
                    row = sc.gettuple()
                    ref = oracle.scan(s, q[0], labels, L, rescore, i)
                    if len(ref["tid"]) < i:
                        assert row is None
                        break
                    assert ((row[0] << 16) | row[1]) == int(ref["tid"][-1]) and row[2] == int(ref["node"][-1])
                    if rescore:
                        assert np.float32(row[3]).view(np.uint32) == ref["dist"][-1].view(np.uint32)
                    st = sc.stats()
                    for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len"):
                        assert st[f] == ref["stats"][f], (f, i, labels, L, rescore, shrink)
            if shrink:
                monkeypatch.delenv("DANN_DEBUG_SHRINK")
        sc.end()


def test_device_buffer_entry_point_matches_host_entry_point(lib):
    """dann_search_batch_device (buffers already in HBM, caller's stream) == dann_search_batch."""
    import torch
    s = build_case(2000, 128, COSINE, seed=91, labels=True, R=24, L_build=50)
    q = _queries(s, 40, 17)
    with lib.DiskAnnIndex(s) as idx:
        dev = buffer_device()
        h = idx.search_batch(q, k=10, search_list_size=60, rescore=25)
        d_q = torch.from_numpy(q).to(dev)
        d_tid = torch.empty((40, 10), dtype=torch.int64, device=dev)
        d_dist = torch.empty((40, 10), dtype=torch.float32, device=dev)
        d_cnt = torch.empty(40, dtype=torch.int32, device=dev)
        d_st = torch.empty((40, 6), dtype=torch.int32, device=dev)
        if emulating():
            idx.search_batch_device(dptr(d_q), 10, 60, 25, dptr(d_tid), dptr(d_dist), dptr(d_cnt), dptr(d_st))
        else:
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                idx.search_batch_device(d_q, 10, 60, 25, d_tid, d_dist, d_cnt, d_st, stream=side.cuda_stream)
            side.synchronize()
        assert np.array_equal(d_tid.cpu().numpy().view(np.uint64), h["tid"])
        assert np.array_equal(d_dist.cpu().numpy().view(np.uint32), h["dist"].view(np.uint32))
        assert np.array_equal(d_cnt.cpu().numpy().view(np.uint32), h["count"])
        assert np.array_equal(d_st.cpu().numpy().view(np.uint32)[:, 0], h["stats"]["visits"])
        # keyed variant: labels must be sorted + dedup per query on the device path
        keys = [[3, 9] for _ in range(40)]
        hk = idx.search_batch(q, labels=keys, k=10, search_list_size=60, rescore=25)
        d_lab = torch.tensor([3, 9] * 40, dtype=torch.int16, device=dev)
        d_off = torch.arange(0, 82, 2, dtype=torch.int32, device=dev)
        idx.search_batch_device(dptr(d_q), 10, 60, 25, dptr(d_tid), dptr(d_dist), dptr(d_cnt), dptr(d_st),
                                d_labels=dptr(d_lab), d_label_off=dptr(d_off))
        if not emulating():
            torch.cuda.synchronize()
        assert np.array_equal(d_tid.cpu().numpy().view(np.uint64), hk["tid"])
