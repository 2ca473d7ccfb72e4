"""The search-kernel SOURCES (pgvectorscale_b200/csrc/dann_search.cuh, dann_search2.cuh, dann_heap.cuh) compiled by
g++ and run on the CPU under a SIMT emulator (tests/simt/simt_emu.h: one fiber per CUDA thread, warp collectives,
named barriers), with the product's own workspace plan (dann_plan.h), compared with the oracle: same approximate
stream (node ids in consume order) and same counters.

This is a LOGIC check that needs no GPU; it does not replace the `-m gpu` parity tests (no memory-model or timing
effects are modelled) and nothing in the product can reach it."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
from conftest import build_case  # noqa: E402
from oracle import fixtures, oracle  # noqa: E402

COSINE, L2, IP = 0, 1, 2


@pytest.fixture(scope="module")
def emu():
    import emu as m
    m.lib()
    return m


def qcodes(s, queries):
    out = []
    for x in queries:
        v = np.array(x[:s.dim_index], np.float32)
        if s.distance_type == COSINE:
            v = oracle.preprocess_cosine(v)
        out.append(oracle.quantize(v, s.bits, s.mean, s.m2, s.count))
    return np.stack(out)


def check(emu, s, queries, L, c_target, labels=None, **kw):
    norm = None if labels is None else [sorted(set(int(x) for x in ls)) for ls in labels]
    streams, st, info = emu.search(s, qcodes(s, queries), L, c_target, labels=norm, **kw)
    for b, q in enumerate(queries):
        r = oracle.scan(s, q, None if labels is None else labels[b], L, 0, c_target)
        assert streams[b].tolist() == r["node"].tolist(), (b, info)
        for f in ("visits", "d_quantized", "candidates", "stream_len"):
            assert st[b][f] == r["stats"][f], (b, f, info)
        assert st[b]["status"] == 0
    return info


@pytest.mark.parametrize("single_warp", [False, True])
@pytest.mark.parametrize("dist,bits,dim", [(COSINE, 2, 96), (L2, 1, 200), (IP, 2, 40)])
def test_emulated_kernels_equal_oracle(emu, single_warp, dist, bits, dim):
    s = build_case(500, dim, dist, bits=bits, seed=11 + dim, kind="normal", R=20, L_build=40, deleted_every=9)
    q = fixtures.gen_vectors(4, dim, 5, "normal")
    info = check(emu, s, q, 25, 30, single_warp=single_warp)
    assert info["pairs"] == (0 if single_warp else 1)


def test_emulated_reference_shape_768d_2bit(emu):
    """768 dimensions x 2 bits = 24 words: the NCH=3, G=4 code mapping the benchmarks run."""
    s = build_case(300, 768, COSINE, seed=2, kind="normal", R=32, L_build=48)
    q = fixtures.gen_vectors(2, 768, 8, "normal")
    info = check(emu, s, q, 20, 29)
    assert (info["nch"], info["G"]) == (3, 4)


@pytest.mark.parametrize("entry", [0, 1, 2])
@pytest.mark.parametrize("single_warp", [False, True])
def test_emulated_heap_entry_layouts(emu, entry, single_warp):
    s = build_case(400, 64, L2, seed=21, kind="normal", R=16, L_build=32)
    q = fixtures.gen_vectors(3, 64, 6, "normal")
    info = check(emu, s, q, 30, 40, single_warp=single_warp, env={"DANN_SEARCH_ENTRY": entry})
    assert info["entry"] == entry


@pytest.mark.parametrize("hs", [8, 64, 256])
@pytest.mark.parametrize("single_warp", [False, True])
def test_emulated_heap_tail_in_global_memory(emu, hs, single_warp):
    """A tiny shared-memory heap top forces every SplitStore path: leaves in the tail, parents in the tail, the pop's
    descent crossing from shared to global memory."""
    s = build_case(500, 64, COSINE, seed=5, kind="normal", R=24, L_build=48)
    q = fixtures.gen_vectors(3, 64, 7, "normal")
    info = check(emu, s, q, 40, 50, single_warp=single_warp, env={"DANN_SEARCH_HS": hs})
    assert info["hs"] == hs


@pytest.mark.parametrize("single_warp", [False, True])
def test_emulated_label_filtered_scan_and_hash_set(emu, single_warp):
    s = build_case(600, 48, L2, seed=31, kind="normal", R=24, L_build=48, labels=True, deleted_every=13)
    q = fixtures.gen_vectors(4, 48, 3, "normal")
    labs = [[3], [7, 1, 7], [], [16, 2, 9, 4]]
    check(emu, s, q, 30, 25, labels=labs, single_warp=single_warp)
    info = check(emu, s, q, 30, 25, labels=labs, single_warp=single_warp, env={"DANN_SEARCH_BITMAP": 0})
    assert info["bitmap_words"] == 0


@pytest.mark.parametrize("single_warp", [False, True])
def test_emulated_workspace_growth_retries(emu, single_warp):
    s = build_case(500, 64, COSINE, seed=41, kind="normal", R=24, L_build=48)
    q = fixtures.gen_vectors(3, 64, 4, "normal")
    info = check(emu, s, q, 40, 60, single_warp=single_warp, env={"DANN_DEBUG_SHRINK": 16})
    assert info["retries"] >= 1


def test_emulated_several_pairs_per_block_and_stream_exhaustion(emu):
    """sm_count=1 packs 7 query slots (14 warps, 7 named barriers) into one block; c_target > n drains the graph."""
    s = build_case(150, 32, L2, seed=51, kind="normal", R=12, L_build=24, deleted_every=5)
    q = fixtures.gen_vectors(9, 32, 2, "normal")
    info = check(emu, s, q, 10, 400, sm_count=1)
    assert info["grid"] == 1 and info["W"] == 7


def test_emulated_lists_with_repeated_ids(emu):
    s = build_case(300, 32, L2, seed=61, kind="normal", R=16, L_build=32)
    nb = s.nbrs.copy()
    nb[:, 5] = nb[:, 1]                 # the same id twice in every list: the per-list dedupe must keep list order
    nb[::3, 9] = nb[::3, 0]
    s.nbrs = nb
    q = fixtures.gen_vectors(3, 32, 12, "normal")
    check(emu, s, q, 20, 30)
    check(emu, s, q, 20, 30, single_warp=True)


# ---- the heap warp's engine alone, differential against the Python clone of Rust's BinaryHeap -----------------
def _model(ops):
    import pyref
    h = pyref.RustBinaryHeap(lambda a, b: b[0] <= a[0])      # Reverse<(key, seq)> with ties Equal
    seq, pops = 0, []
    for o in ops:
        if o[0] == "push":
            for k in o[1]:
                h.push((int(k), seq))
                seq += 1
        elif len(h):
            pops.append(h.pop()[1])
    return list(h.data), pops


def _random_script(rng, nops, key_range, pop_share, first=None):
    ops = [("push", first)] if first else []
    for _ in range(nops):
        if rng.random() < pop_share:
            ops.append(("pop",))
        else:
            ops.append(("push", [int(x) for x in rng.integers(0, key_range, size=int(rng.integers(1, 65)))]))
    return ops


@pytest.mark.parametrize("hv", [0, 1, 2])
@pytest.mark.parametrize("entry,key_range", [(0, 4), (0, 1500), (1, 40), (2, 3)])
@pytest.mark.parametrize("hs", [8, 256, 16384])
def test_emulated_heap_engine_equals_rust_heap_model(emu, monkeypatch, hv, entry, key_range, hs):
    """Random pages of pushes and pops with few distinct keys (ties everywhere, elements racing to the root, pages
    that straddle a leaf-level change): the heap ARRAY after the script and every popped element must equal the
    sequential std algorithm's, for the shared-memory part and the global tail, under all three lane schedules."""
    rng = np.random.default_rng(1000 * hv + 100 * entry + key_range + hs)
    for sched in (0, 1, 2):
        monkeypatch.setenv("SIMT_SCHED", str(sched))
        ops = _random_script(rng, 120, key_range, 0.45)
        heap, pops = emu.heap_script(ops, entry=entry, hv=hv, hs=hs)
        want_heap, want_pops = _model(ops)
        assert heap == want_heap
        assert pops[:len(want_pops)].tolist() == want_pops


@pytest.mark.parametrize("hv", [0, 1])
def test_emulated_heap_engine_root_handover_at_leaf_level_change(emu, monkeypatch, hv):
    """The push that fills slot 2^d - 1 rises to the root, and the next page starts a new leaf level: the root slot
    changes owner lane exactly there (engine v2 needs its write-back / sync / reload)."""
    for sched in (0, 1, 2):
        monkeypatch.setenv("SIMT_SCHED", str(sched))
        for d in (7, 8, 10):
            n0 = (1 << d) - 1 - 40
            ops = [("push", [1000] * 64) for _ in range(n0 // 64)] + [("push", [1000] * (n0 % 64))]
            # one page: 39 large keys, then a new minimum at slot 2^d - 1, then smaller and smaller keys beyond it
            ops.append(("push", [900] * 39 + [5] + [4, 3, 3, 2, 900, 1, 0, 0]))
            ops.append(("pop",))
            ops.append(("push", [0, 7, 0]))
            heap, pops = emu.heap_script([o for o in ops if o[0] == "pop" or len(o[1])], entry=0, hv=hv, hs=4096)
            want_heap, want_pops = _model([o for o in ops if o[0] == "pop" or len(o[1])])
            assert heap == want_heap and pops[:len(want_pops)].tolist() == want_pops


# ---- plain storage layout: SearchWarp<Ent64, 1, PLAIN=1> (f32 keys, exact distances inside the beam search) ----
def qindex(s, queries):
    out = []
    for x in queries:
        v = np.array(x[:s.dim_index], np.float32)
        if s.distance_type == COSINE:
            v = oracle.preprocess_cosine(v)
        out.append(v)
    return np.stack(out)


@pytest.mark.parametrize("dist,dim,dim_index", [(COSINE, 64, None), (L2, 64, None), (COSINE, 96, 40), (L2, 70, 38),
                                                (L2, 33, None), (COSINE, 20, 7)])
def test_emulated_plain_storage_scan_equals_oracle(emu, monkeypatch, dist, dim, dim_index):
    """Vector widths with and without whole 32-element strides, multiples of 4 or not (vector vs scalar loads, the
    scalar tail), truncated index slices; deleted tuples; all three lane schedules."""
    s = fixtures.to_plain(build_case(500, dim, dist, seed=3 + dim, kind="normal", R=24, L_build=48, deleted_every=11,
                                     dim_index=dim_index))
    q = fixtures.gen_vectors(4, dim, 9, "normal")
    for sched in (0, 1, 2):
        monkeypatch.setenv("SIMT_SCHED", str(sched))
        streams, st, info = emu.search(s, None, 30, 40, q_index=qindex(s, q))
        assert info["entry"] == 2 and info["pairs"] == 0
        for b in range(len(q)):
            r = oracle.scan(s, q[b], None, 30, 0, 40)
            assert streams[b].tolist() == r["node"].tolist()
            assert st[b]["d_quantized"] == 0 and st[b]["status"] == 0
            # rescore=0 in the oracle call: its d_full is exactly the beam search's comparisons
            for f in ("visits", "candidates", "d_full", "stream_len"):
                assert st[b][f] == r["stats"][f], f


def test_emulated_plain_storage_tail_and_retries(emu):
    s = fixtures.to_plain(build_case(900, 48, L2, seed=15, kind="normal", R=32, L_build=64))
    q = fixtures.gen_vectors(3, 48, 4, "normal")
    streams, st, info = emu.search(s, None, 60, 120, q_index=qindex(s, q),
                                   env={"DANN_SEARCH_HS": 32, "DANN_DEBUG_SHRINK": 16, "DANN_SEARCH_BITMAP": 0})
    assert info["retries"] >= 1 and info["hs"] == 32 and info["bitmap_words"] == 0
    for b in range(len(q)):
        assert streams[b].tolist() == oracle.scan(s, q[b], None, 60, 0, 120)["node"].tolist()


# ---- a whole dann_search_batch call replayed on the CPU: prepare -> search -> rerank kernels under emulation ------
def _batch_under_emulation(emu, s, q, k, L, rescore, labels=None, plain=False, env=None):
    """Mirrors search_batch_device_locked (diskann_b200.cu): same kernels, same order, same stream/rerank sizes."""
    if plain and s.dim == s.dim_index:
        rescore = 0                                         # scan.rs:392-403
    c_target = k if rescore == 0 else rescore + k - 1       # scan.rs:255-305
    if plain:
        q_full, q_index = emu.prepare_plain(s, q)
        streams, st, _ = emu.search(s, None, L, c_target, q_index=q_index, env=env)
    else:
        q_full, q_codes = emu.prepare(s, q)
        norm = None if labels is None else [sorted(set(int(x) for x in ls)) for ls in labels]
        streams, st, _ = emu.search(s, q_codes, L, c_target, labels=norm, env=env)
    return emu.rerank(s, q_full, streams, c_target, k, rescore, stats=st, plain=plain)


def _assert_batch_equals_oracle(s, q, g, k, L, rescore, labels=None, dist_bits=True):
    for b in range(len(q)):
        r = oracle.scan(s, q[b], None if labels is None else labels[b], L, rescore, k)
        n = len(r["tid"])
        assert int(g["count"][b]) == n
        assert g["tid"][b, :n].tolist() == r["tid"].tolist()
        assert g["node"][b, :n].tolist() == r["node"].tolist()
        if dist_bits:
            assert g["dist"][b, :n].view(np.uint32).tolist() == r["dist"].view(np.uint32).tolist()
        for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len"):
            assert g["stats"][b][f] == r["stats"][f], f


@pytest.mark.parametrize("dist,bits,dim,dim_index", [(COSINE, 2, 96, None), (L2, 1, 70, None), (IP, 2, 64, 40),
                                                     (COSINE, 2, 48, 33)])
def test_emulated_whole_batch_call_equals_oracle(emu, dist, bits, dim, dim_index):
    """Row ids, rerank distance BITS and every counter of a full batch call, with no GPU in the box."""
    s = build_case(600, dim, dist, bits=bits, seed=7 + dim, kind="normal", R=24, L_build=48, deleted_every=9,
                   dim_index=dim_index)
    q = fixtures.gen_vectors(4, dim, 21, "normal")
    for (k, L, rescore) in ((10, 30, 20), (5, 10, 0), (12, 40, 64)):
        g = _batch_under_emulation(emu, s, q, k, L, rescore)
        _assert_batch_equals_oracle(s, q, g, k, L, rescore, dist_bits=rescore > 0)


def test_emulated_whole_batch_call_with_label_keys(emu):
    s = build_case(700, 48, L2, seed=31, kind="normal", R=24, L_build=48, labels=True, deleted_every=13)
    q = fixtures.gen_vectors(4, 48, 3, "normal")
    labs = [[3], [7, 1, 7], [], [16, 2, 9, 4]]
    g = _batch_under_emulation(emu, s, q, 10, 30, 25, labels=labs)
    _assert_batch_equals_oracle(s, q, g, 10, 30, 25, labels=labs)


@pytest.mark.parametrize("dist,dim,dim_index", [(COSINE, 64, None), (L2, 70, 38), (COSINE, 96, 40)])
def test_emulated_whole_plain_storage_batch_call_equals_oracle(emu, dist, dim, dim_index):
    """Plain layout end to end: plain prepare kernel, PLAIN search kernel, rerank only when the index slice is
    shorter than the heap vector, counters d_quantized = 0 / d_full = candidates + reranked."""
    s = fixtures.to_plain(build_case(500, dim, dist, seed=5 + dim, kind="normal", R=24, L_build=48, deleted_every=11,
                                     dim_index=dim_index))
    q = fixtures.gen_vectors(4, dim, 17, "normal")
    for (k, L, rescore) in ((10, 30, 20), (8, 15, 0)):
        g = _batch_under_emulation(emu, s, q, k, L, rescore, plain=True)
        _assert_batch_equals_oracle(s, q, g, k, L, rescore, dist_bits=rescore > 0 and s.dim != s.dim_index)


@pytest.mark.parametrize("name", ["cos768b2", "l2_768b1", "ip100lab"])
def test_sbq_golden_vectors_emulated_kernels(emu, name):
    """tests/golden/scan_golden.npz (the vectors the GPU parity test uses) reproduced by the emulated kernels."""
    from golden.make_golden import CASES, make_case
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_golden.npz"))
    s, q, lab, L, rescore, k = make_case(name)
    g = _batch_under_emulation(emu, s, q, k, L, rescore, labels=lab)
    assert np.array_equal(g["count"], z[f"{name}/count"])
    assert np.array_equal(g["tid"], z[f"{name}/tid"])
    if rescore:
        for b in range(len(q)):
            n = int(g["count"][b])
            assert g["dist"][b, :n].view(np.uint32).tolist() == z[f"{name}/dist_bits"][b, :n].tolist()
    assert [x["visits"] for x in g["stats"]] == z[f"{name}/visits"].tolist()
    assert [x["d_quantized"] for x in g["stats"]] == z[f"{name}/d_quantized"].tolist()


@pytest.mark.parametrize("name", ["plain_cos128", "plain_l2_96x40", "plain_cos70x38"])
def test_plain_golden_vectors_oracle_and_emulated_kernels(emu, name):
    """tests/golden/plain_golden.npz: the oracle still gives the frozen answers, and so do the kernels under emulation."""
    from golden.make_plain_golden import CASES, make_case, run_case
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "plain_golden.npz"))
    now = run_case(name)
    for key, val in now.items():
        assert np.array_equal(z[f"{name}/{key}"], val), key
    s, q, L, rescore, k = make_case(name)
    g = _batch_under_emulation(emu, s, q, k, L, rescore, plain=True)
    resorts = rescore > 0 and s.dim != s.dim_index
    for b in range(len(q)):
        n = int(z[f"{name}/count"][b])
        assert int(g["count"][b]) == n
        assert g["tid"][b, :n].tolist() == z[f"{name}/tid"][b, :n].tolist()
        if resorts:
            assert g["dist"][b, :n].view(np.uint32).tolist() == z[f"{name}/dist_bits"][b, :n].tolist()
        assert g["stats"][b]["visits"] == int(z[f"{name}/visits"][b])
        assert g["stats"][b]["d_full"] == int(z[f"{name}/d_full"][b])


# ---- the workspace plan itself (host logic of diskann_b200.cu, shared through dann_plan.h) ----------------
def test_plan_benchmark_shape_is_one_wave_of_seven_pairs(emu):
    p = emu.plan(n=1_000_000, R=64, words=24, nq=1024, L=150, c_target=259)
    assert p["pairs"] == 1 and p["W"] == 7 and p["grid"] == 147          # ceil(1024 / 7) blocks, one per SM
    assert p["entry"] == 0 and p["esize"] == 4                            # 1536-bit codes, < 2M candidates: 4-byte entries
    assert p["bitmap_words"] == (1_000_000 + 127) // 128 * 4
    assert p["per_warp"] * p["W"] <= 232448 - 1024
    assert p["hs"] >= 2048 and p["hs"] % 4 == 0
    assert p["cand_cap"] % 1024 == 0 and p["cand_cap"] >= (150 + 259) * 64


def test_plan_scales_with_growth_and_keys(emu):
    a = emu.plan(n=50_000_000, R=64, words=24, nq=4096, L=1000, c_target=1009)
    assert a["bitmap_words"] == 0 and a["hash_cap"] >= 2 * a["cand_cap"]   # > 16M nodes: CAS hash set
    b = emu.plan(n=50_000_000, R=64, words=24, nq=4096, L=1000, c_target=1009, grow=2, keyed=True)
    assert b["cand_cap"] >= 2 * a["cand_cap"] - 1024 and b["vcap"] > 2 * a["vcap"] - 8
    assert b["hash_cap"] >= 4 * b["cand_cap"]
    c = emu.plan(n=1000, R=100, words=24, nq=10, L=100, c_target=59)        # R > 64: single-warp kernel
    assert c["pairs"] == 0
    with pytest.raises(RuntimeError, match="2\\^30"):
        emu.plan(n=1000, R=64, words=24, nq=1, L=10000, c_target=1009, grow=1 << 12)


# ---- lean warp-per-query kernel (dann_search3.cuh): the product's default for batch searches ---------------------
@pytest.mark.parametrize("dist,bits,dim", [(COSINE, 2, 96), (L2, 1, 200), (IP, 2, 40)])
def test_emulated_lean_kernel_equals_oracle(emu, dist, bits, dim):
    s = build_case(500, dim, dist, bits=bits, seed=11 + dim, kind="normal", R=20, L_build=40, deleted_every=9)
    q = fixtures.gen_vectors(4, dim, 5, "normal")
    info = check(emu, s, q, 25, 30, kernel="lean")
    assert info["lean"] == 1 and info["pairs"] == 0


def test_emulated_lean_reference_shape_768d_2bit(emu):
    s = build_case(300, 768, COSINE, seed=2, kind="normal", R=32, L_build=48)
    q = fixtures.gen_vectors(2, 768, 8, "normal")
    info = check(emu, s, q, 20, 29, kernel="lean")
    assert (info["nch"], info["G"], info["lean"], info["entry"]) == (3, 4, 1, 0)


@pytest.mark.parametrize("flags", [2, 0, 3])
@pytest.mark.parametrize("entry", [0, 2])
@pytest.mark.parametrize("bitmap", [0, 1])
@pytest.mark.parametrize("hs", [None, 4, 16, 128, 1024])
def test_emulated_lean_long_scans_entries_sets_and_tails(emu, entry, bitmap, hs, flags):
    """Long scans: the heap crosses several leaf levels, the pop's four-level rounds straddle the shared/global split,
    the visited ring wraps; both inserted-set flavours (node-carrying vs hash-slot-carrying 4-byte entries)."""
    s = build_case(2500, 64, COSINE, seed=71, kind="normal", R=32, L_build=64, deleted_every=17)
    q = fixtures.gen_vectors(3, 64, 13, "normal")
    # flags: bit 2 = four-level pop rounds (else lane 0 walks the hole down), bit 1 = staged pushes even when the page is all in shared memory
    env = {"DANN_SEARCH_ENTRY": entry, "DANN_SEARCH_BITMAP": bitmap, "DANN_HV_FLAGS": flags}
    if hs is not None:
        env["DANN_SEARCH_HS"] = hs
    info = check(emu, s, q, 80, 150, env=env, kernel="lean")
    assert info["lean"] == 1 and info["entry"] == entry and (info["bitmap_words"] != 0) == bool(bitmap)


@pytest.mark.parametrize("L", [1, 2, 31, 32, 33, 64, 200, 1100])
def test_emulated_lean_visited_ring_list_sizes(emu, L):
    s = build_case(1500, 48, L2, seed=91, kind="normal", R=24, L_build=48, deleted_every=11)
    q = fixtures.gen_vectors(2, 48, 17, "normal")
    check(emu, s, q, L, 40, kernel="lean")
    check(emu, s, q, L, 40, kernel="lean", env={"DANN_SEARCH_BITMAP": 0, "DANN_SEARCH_HS": 32})


def test_emulated_lean_labels_retries_and_packed_block(emu):
    s = build_case(800, 48, L2, seed=81, kind="normal", R=24, L_build=48, labels=True, deleted_every=13)
    q = fixtures.gen_vectors(6, 48, 3, "normal")
    labs = [[3], [7, 1, 7], [], [16, 2, 9, 4], [5, 5], [12]]
    check(emu, s, q, 30, 25, labels=labs, kernel="lean")
    check(emu, s, q, 100, 59, labels=labs, kernel="lean", env={"DANN_HV_FLAGS": 0})   # heap runs empty: pop with 1-3 entries
    info = check(emu, s, q, 30, 25, labels=labs, kernel="lean", env={"DANN_SEARCH_BITMAP": 0})
    assert info["bitmap_words"] == 0
    info = check(emu, s, q, 40, 60, labels=labs, kernel="lean", env={"DANN_DEBUG_SHRINK": 16}, sm_count=1)
    assert info["retries"] >= 1 and info["grid"] == 1 and info["lean"] == 1


def test_emulated_lean_repeated_ids_and_stream_exhaustion(emu):
    s = build_case(300, 32, L2, seed=61, kind="normal", R=16, L_build=32, deleted_every=5)
    nb = s.nbrs.copy()
    nb[:, 5] = nb[:, 1]
    nb[::3, 9] = nb[::3, 0]
    s.nbrs = nb
    q = fixtures.gen_vectors(3, 32, 12, "normal")
    check(emu, s, q, 20, 30, kernel="lean")
    check(emu, s, q, 10, 400, kernel="lean", env={"DANN_SEARCH_BITMAP": 0})


def test_emulated_lean_at_the_benchmark_operating_point(emu):
    s = build_case(6000, 768, COSINE, seed=3, kind="normal", R=50, L_build=100)
    q = fixtures.gen_vectors(2, 768, 9, "normal")
    info = check(emu, s, q, 150, 259, kernel="lean", sm_count=1)
    assert info["lean"] == 1
    info = check(emu, s, q, 150, 259, kernel="lean", env={"DANN_SEARCH_BITMAP": 0, "DANN_SEARCH_HS": 256})
    assert info["bitmap_words"] == 0
