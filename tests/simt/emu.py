"""ctypes front end of the CPU SIMT-emulation harness (tests/simt/emu_search.cpp).  Test infrastructure only."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_emu  # noqa: E402
from pgvectorscale_b200.diskann import _SnapshotDesc, _QueryStats  # noqa: E402  (struct layouts of include/diskann_b200.h)


class EmuInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("retries", "entry", "W", "hs", "pairs", "grid", "cand_cap", "vcap",
                                          "bitmap_words", "nch", "G", "hv", "lean", "maxw", "hash_cap")] + \
        [("switches", C.c_uint64), ("coll_even", C.c_uint64), ("coll_odd", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.environ.get("DANN_EMU_LIB") or build_emu.build())
        _lib.emu_last_error.restype = C.c_char_p
        _lib.emu_search.restype = C.c_int
    return _lib


def _desc(s):
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return C.c_void_p(a.ctypes.data)

    d = _SnapshotDesc()
    d.n, d.dim, d.dim_index, d.bits, d.words, d.R = s.n, s.dim, s.dim_index, s.bits, s.words, s.R
    d.distance_type = int(s.distance_type)
    d.has_labels = int(bool(s.has_labels))
    d.count = int(s.count)
    d.mean = arr(s.mean, np.float32)
    d.m2 = arr(s.m2, np.float32)
    d.codes = arr(s.codes, np.uint64)
    d.nbrs = arr(s.nbrs, np.uint32)
    d.heap_tid = arr(s.heap_tid, np.uint64)
    d.vectors = arr(s.vectors, np.float32)
    d.start_default = int(s.start_default)
    d.n_start_labels = 0 if s.start_labels is None else len(s.start_labels)
    d.start_labels = arr(s.start_labels, np.int16)
    d.start_label_nodes = arr(s.start_label_nodes, np.uint32)
    d.label_off = arr(s.label_off, np.uint32)
    d.labels = arr(s.labels, np.int16)
    return d, keep


def search(s, q_codes, L, c_target, labels=None, single_warp=False, sm_count=148, smem_optin=232448, env=None,
           q_index=None, kernel=None):
    """Approximate streams of B prepared queries through the emulated search kernel.
    labels: None, or a list (one entry per query) of sorted, de-duplicated label lists.
    env: DANN_* test knobs applied around the call (the plan reads them with getenv).
    -> (streams: list of np.uint32 arrays, stats: list of dicts, info dict)"""
    plain = q_index is not None   # plain storage layout: the kernel reads s.index_vectors / q_index, no codes
    if plain:
        q_index = np.ascontiguousarray(q_index, dtype=np.float32).reshape(-1, s.dim_index)
        iv = np.ascontiguousarray(s.index_vectors, dtype=np.float32)
        B = q_index.shape[0]
        q_codes = np.zeros((B, max(s.words, 1)), np.uint64)
    else:
        q_codes = np.ascontiguousarray(q_codes, dtype=np.uint64).reshape(-1, s.words)
        B = q_codes.shape[0]
    d, keep = _desc(s)
    qlab = qoff = None
    if labels is not None:
        off = np.zeros(B + 1, np.int32)
        flat = []
        for b, ls in enumerate(labels):
            flat.extend(int(x) for x in ls)
            off[b + 1] = len(flat)
        qlab = np.array(flat if flat else [0], np.int16)
        qoff = off
    stream = np.full((B, max(c_target, 1)), 0xFFFFFFFF, np.uint32)
    slen = np.zeros(B, np.uint32)
    stats = (_QueryStats * B)()
    info = EmuInfo()
    old = {}
    env = dict(env or {})
    # kernel: "lean" = dann_search3.cuh (the product's default for batch searches), None/"pairs" = the two-warp kernel
    # (or the single-warp one with single_warp=True)
    env.setdefault("DANN_SEARCH_KERNEL", 3 if kernel == "lean" else 2)
    for k, v in env.items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    try:
        rc = lib().emu_search(C.byref(d), C.c_void_p(q_codes.ctypes.data),
                              C.c_void_p(qlab.ctypes.data) if qlab is not None else None,
                              C.c_void_p(qoff.ctypes.data) if qoff is not None else None,
                              C.c_uint32(B), C.c_uint32(L), C.c_uint32(c_target), C.c_int(1 if single_warp else 0),
                              C.c_uint32(sm_count), C.c_uint32(smem_optin), C.c_void_p(stream.ctypes.data),
                              C.c_void_p(slen.ctypes.data), stats, C.byref(info),
                              C.c_void_p(iv.ctypes.data) if plain else None,
                              C.c_void_p(q_index.ctypes.data) if plain else None)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc != 0:
        raise RuntimeError(f"emu_search failed ({rc}): {lib().emu_last_error().decode()}")
    streams = [stream[b, :slen[b]].copy() for b in range(B)]
    st = [{f: getattr(stats[b], f) for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len", "status")}
          for b in range(B)]
    return streams, st, {f: getattr(info, f) for f, _ in EmuInfo._fields_}


PLAN_FIELDS = ("need", "cand_cap", "hash_cap", "vcap", "hs", "W", "grid", "per_warp", "esize", "bitmap_words", "ins_cap",
               "entry", "pairs")


def plan(n, R, words, nq, L, c_target, grow=1, keyed=False, force_single=False, sm_count=148, smem_optin=232448):
    """dann_make_plan (pgvectorscale_b200/csrc/dann_plan.h) -> dict; raises on DANN_ERR_CAPACITY."""
    out = (C.c_uint32 * 13)()
    rc = lib().emu_plan(C.c_uint32(n), C.c_uint32(R), C.c_uint32(words), C.c_uint32(nq), C.c_uint32(L), C.c_uint32(c_target),
                        C.c_uint32(grow), C.c_int(int(keyed)), C.c_int(int(force_single)), C.c_uint32(sm_count),
                        C.c_uint32(smem_optin), out)
    if rc != 0:
        raise RuntimeError(f"plan failed ({rc}): {lib().emu_last_error().decode()}")
    return dict(zip(PLAN_FIELDS, [int(x) for x in out]))


KSHIFT = {0: 21, 1: 16, 2: 32}


def heap_script(ops, entry=0, hv=0, hs=64, cap=1 << 16):
    """ops: list of ("push", [keys...]) (<= 64 keys: one page) or ("pop",).  Runs the heap warp's engine under the
    emulator.  -> (heap as list of (key, seq) for slots 1..len, popped sequence numbers)"""
    kinds = np.array([0 if o[0] == "push" else 1 for o in ops], np.uint32)
    n = np.array([len(o[1]) if o[0] == "push" else 0 for o in ops], np.uint32)
    keys = np.array([k for o in ops if o[0] == "push" for k in o[1]] or [0], np.uint32)
    out = np.zeros(cap, np.uint64)
    olen = np.zeros(1, np.uint32)
    pops = np.zeros(max(1, int((kinds == 1).sum())), np.uint32)
    rc = lib().emu_heap_script(C.c_int(entry), C.c_int(hv), C.c_void_p(kinds.ctypes.data), C.c_void_p(n.ctypes.data),
                               C.c_uint32(len(ops)), C.c_void_p(keys.ctypes.data), C.c_uint32(hs), C.c_uint32(cap),
                               C.c_void_p(out.ctypes.data), C.c_void_p(olen.ctypes.data), C.c_void_p(pops.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"emu_heap_script failed ({rc}): {lib().emu_last_error().decode()}")
    sh = KSHIFT[entry] if (hv == 0 or entry != 1) else 32     # the lean kernel has no 16-bit-key layout: Ent64 there
    heap = [(int(e) >> sh, int(e) & ((1 << sh) - 1)) for e in out[:int(olen[0])]]
    return heap, pops


def prepare(s, queries):
    """dann_prepare_kernel under emulation -> (q_full [B, dim], q_codes [B, words])"""
    queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, s.dim)
    B = queries.shape[0]
    d, keep = _desc(s)
    q_full = np.zeros((B, s.dim), np.float32)
    q_codes = np.zeros((B, s.words), np.uint64)
    rc = lib().emu_prepare(C.byref(d), C.c_void_p(queries.ctypes.data), C.c_int(B), C.c_void_p(q_full.ctypes.data),
                           C.c_void_p(q_codes.ctypes.data))
    assert rc == 0
    return q_full, q_codes


def prepare_plain(s, queries):
    """dann_prepare_plain_kernel under emulation -> (q_full [B, dim], q_index [B, dim_index])"""
    queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, s.dim)
    B = queries.shape[0]
    q_full = np.zeros((B, s.dim), np.float32)
    q_index = np.zeros((B, s.dim_index), np.float32)
    rc = lib().emu_prepare_plain(C.c_uint32(s.dim), C.c_uint32(s.dim_index), C.c_int(int(s.distance_type == 0)),
                                 C.c_void_p(queries.ctypes.data), C.c_int(B), C.c_void_p(q_full.ctypes.data),
                                 C.c_void_p(q_index.ctypes.data))
    assert rc == 0
    return q_full, q_index


def rerank(s, q_full, streams, c_target, k, rescore, stats=None, plain=False):
    """dann_rerank_kernel (after dann_normalize_rows_kernel for cosine) under emulation.
    -> dict(tid [B,k], dist [B,k], node [B,k], count [B], d_full [B])"""
    B = len(streams)
    d, keep = _desc(s)
    q_full = np.ascontiguousarray(q_full, dtype=np.float32)
    st = np.full((B, max(c_target, 1)), 0xFFFFFFFF, np.uint32)
    sl = np.zeros(B, np.uint32)
    for b, x in enumerate(streams):
        st[b, :len(x)] = x
        sl[b] = len(x)
    tid = np.zeros((B, k), np.uint64)
    dist = np.zeros((B, k), np.float32)
    node = np.zeros((B, k), np.uint32)
    count = np.zeros(B, np.uint32)
    qs = (_QueryStats * B)()
    if stats is not None:
        for b in range(B):
            for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len", "status"):
                setattr(qs[b], f, stats[b][f])
    rc = lib().emu_rerank(C.byref(d), C.c_void_p(q_full.ctypes.data), C.c_void_p(st.ctypes.data), C.c_void_p(sl.ctypes.data),
                          C.c_int(B), C.c_uint32(c_target), C.c_uint32(k), C.c_uint32(rescore), C.c_void_p(tid.ctypes.data),
                          C.c_void_p(dist.ctypes.data), C.c_void_p(node.ctypes.data), C.c_void_p(count.ctypes.data), qs)
    assert rc == 0
    if plain and rescore > 0:
        assert lib().emu_plain_stats(qs, C.c_int(B)) == 0
    return dict(tid=tid, dist=dist, node=node, count=count,
                stats=[{f: getattr(qs[b], f) for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len")}
                       for b in range(B)])
