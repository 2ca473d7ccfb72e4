"""the relation-file reader (dann_pg_*, csrc/dann_pgreader.h) against relation images written by
tests/pgpages.py with the reference's own write rules (util/page.rs, util/chain.rs, util/tape.rs, meta_page.rs:360-384,
sbq/node.rs:26-42).  Host-only entry points of the C-ABI library: they run on the CPU box.  Parity here is
self-consistent (writer and reader are both restatements; nothing in the sandbox can produce a file with the reference
itself) - the reader's header says so."""
import os
import struct

import numpy as np
import pytest

from conftest import build_case
from pgvectorscale_b200.snapshot import COSINE, L2
import pgpages


@pytest.fixture(scope="module")
def pg(lib_built):
    from pgvectorscale_b200 import pgreader
    return pgreader


def _same_snapshot(a, b):
    for f in ("n", "dim", "dim_index", "bits", "words", "R", "distance_type", "count", "start_default"):
        assert getattr(a, f) == getattr(b, f), f
    assert bool(a.has_labels) == bool(b.has_labels)
    for f in ("codes", "nbrs", "heap_tid", "mean", "m2", "start_labels", "start_label_nodes", "label_off", "labels"):
        x, y = getattr(a, f), getattr(b, f)
        if x is None or y is None:
            assert (x is None or len(x) == 0) and (y is None or len(y) == 0), f
        else:
            assert np.array_equal(np.asarray(x), np.asarray(y)), f


@pytest.mark.parametrize("labels", [False, True])
@pytest.mark.parametrize("order", [(0, 1, 2, 3), (3, 0, 2, 1), (1, 2, 3, 0)])
def test_extracted_snapshot_equals_the_one_written(pg, tmp_path, labels, order):
    s = build_case(700, 96, COSINE, seed=5, R=24, L_build=48, labels=labels, deleted_every=9)
    path = str(tmp_path / "16384")
    meta, tids, w = pgpages.write_index(s, path, field_order=order, dead_every=7)
    with pg.PgRelation(path) as rel:
        info = rel.info()
        assert info["meta_magic"] == pgpages.TSV_MAGIC and info["meta_version"] == pgpages.TSV_VERSION
        assert info["node_items"] == s.n and info["foreign_pages"] == 0 and info["new_pages"] == 0
        assert info["pages_by_type"]["Meta"] == 1 and info["pages_by_type"]["SbqNode"] > 10
        assert info["nblocks"] == len(w.pages)
        got, index_tid, fp, layout = rel.extract_sbq(pg.PgMeta(**meta))
    assert layout == order and fp == info["fingerprint"]
    assert np.array_equal(index_tid, tids)            # dense ids = (block, offset) order, dead line pointers skipped
    _same_snapshot(got, s)
    assert got.vectors is None                        # heap rows are not in the index relation


def test_scan_of_the_extracted_snapshot_equals_scan_of_the_original(pg, tmp_path):
    """End of the row: the oracle scans the original snapshot and the one that went through the page format."""
    from oracle import fixtures, oracle
    s = build_case(500, 64, L2, seed=8, R=16, L_build=32, labels=True)
    path = str(tmp_path / "rel")
    meta, _, _ = pgpages.write_index(s, path)
    with pg.PgRelation(path) as rel:
        got, _, _, _ = rel.extract_sbq(pg.PgMeta(**meta))
    got.vectors = s.vectors                            # the host fetches heap_tid[i] from the table
    q = fixtures.gen_vectors(6, 64, 3, "normal")
    labs = [[1], [2, 5], [], [7], [3, 3], [9]]
    lo = np.cumsum([0] + [len(x) for x in labs]).astype(np.int32)
    lf = np.array([y for x in labs for y in x], dtype=np.int16)
    a = oracle.scan_batch(s, q, lf, lo, 30, 20, 10)
    b = oracle.scan_batch(got, q, lf, lo, 30, 20, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_chains_across_pages_and_segments(pg, tmp_path, monkeypatch):
    """util/chain.rs tests: payload sizes around one, two and three pages come back byte for byte; a relation split
    into several segment files reads like one."""
    w = pgpages.RelationWriter()
    b0 = w.new_page(pgpages.PT_SBQ_MEANS)
    items = []
    cur = b0
    for size in [1, 14, 8100, 8140, 8192 - 100, 8192 + 100, 2 * 8192 - 100, 2 * 8192 + 100, 3 * 8192 + 7]:
        data = bytes((i * 7 + size) % 256 for i in range(size))
        ip, cur = w.chain_write(pgpages.PT_SBQ_MEANS, cur, data)
        items.append((ip, data))
    path = str(tmp_path / "chain")
    w.save(path, segment_blocks=4)
    monkeypatch.setenv("DANN_PG_RELSEG_BLOCKS", "4")       # RELSEG_SIZE of this "server" (1 GB = 131072 blocks by default)
    assert os.path.exists(path + ".1") and os.path.exists(path + ".2")
    with pg.PgRelation(path) as rel:
        assert rel.info()["nblocks"] == len(w.pages)
        for (blk, off), data in items:
            assert rel.read_chain(blk, off, pgpages.PT_SBQ_MEANS) == data
        with pytest.raises(Exception) as e:
            rel.read_chain(items[0][0][0], items[0][0][1], pgpages.PT_META)     # assert!(page.get_type() == self.page_type)
        assert "PageType" in str(e.value)
        with pytest.raises(Exception):
            rel.read_chain(len(w.pages) + 3, 1)


def test_fingerprint_moves_with_any_page_change(pg, tmp_path):
    s = build_case(300, 48, L2, seed=2, R=12, L_build=24)
    path = str(tmp_path / "rel")
    meta, _, w = pgpages.write_index(s, path)
    with pg.PgRelation(path) as rel:
        fp0 = rel.info()
    # an insert rewrites a neighbour list in place: same bytes layout, the page's LSN moves (GenericXLogFinish)
    blob = bytearray(open(path, "rb").read())
    blk = 3
    lsn_lo = struct.unpack_from("<I", blob, blk * 8192 + 4)[0]
    struct.pack_into("<I", blob, blk * 8192 + 4, lsn_lo + 0x40)
    open(path, "wb").write(blob)
    with pg.PgRelation(path) as rel:
        fp1 = rel.info()
    assert fp1["fingerprint"] != fp0["fingerprint"] and fp1["max_lsn"] >= fp0["max_lsn"] and fp1["nblocks"] == fp0["nblocks"]
    # relation extension (a new, still all-zero page at the end)
    open(path, "ab").write(b"\0" * 8192)
    with pg.PgRelation(path) as rel:
        fp2 = rel.info()
        assert fp2["fingerprint"] != fp1["fingerprint"] and fp2["new_pages"] == 1
        got, _, fp, _ = rel.extract_sbq(pg.PgMeta(**meta))        # a new page holds no nodes and is not an error
        assert got.n == s.n and fp == fp2["fingerprint"]


def _corrupt(path, pos, data):
    blob = bytearray(open(path, "rb").read())
    blob[pos:pos + len(data)] = data
    open(path, "wb").write(blob)


def test_everything_unexpected_is_refused_not_guessed(pg, tmp_path):
    from pgvectorscale_b200.diskann import DiskAnnError
    s = build_case(200, 48, L2, seed=4, R=12, L_build=24)
    path = str(tmp_path / "rel")
    meta, tids, w = pgpages.write_index(s, path)
    good = open(path, "rb").read()

    def expect(*msgs, **meta_over):
        with pg.PgRelation(path) as rel:
            with pytest.raises(DiskAnnError) as e:
                rel.extract_sbq(pg.PgMeta(**{**meta, **meta_over}))
        assert e.value.code in (-7, -1) and any(m in str(e.value) for m in msgs), str(e.value)

    expect("no order of the archived root's fields", num_neighbors=13)        # wrong R: no layout fits
    expect("no order of the archived root's fields", bq_bits=1)               # wrong code width
    expect("start_nodes.default_node", start=(1, 999))
    # a page of another access method in the middle of the relation
    _corrupt(path, 2 * 8192 + 8192 - 8 + 2, b"\x00\x00")
    expect("page id 0xAE24")
    open(path, "wb").write(good)
    # a neighbour pointer to an item that does not exist
    blk, off = int(tids[5]) >> 16, int(tids[5]) & 0xFFFF
    pgv = w.pages[blk]
    lp = struct.unpack_from("<I", pgv.b, 24 + 4 * (off - 1))[0]
    _corrupt(path, blk * 8192 + (lp & 0x7FFF) + s.words * 8, struct.pack("<IHH", blk, 200, 0))
    expect("not a live node item", "no order of the archived root's fields")     # caught by the layout sample or by the full pass
    open(path, "wb").write(good)
    # truncated file: not a whole number of pages
    open(path, "wb").write(good[:-100])
    with pytest.raises(DiskAnnError) as e:
        pg.PgRelation(path)
    assert e.value.code == -7
    # pd_upper beyond pd_special
    open(path, "wb").write(good)
    _corrupt(path, 1 * 8192 + 14, struct.pack("<H", 8190))
    expect("pd_lower / pd_upper / pd_special")


def test_meta_header_is_recognised_in_either_field_order(pg, tmp_path):
    w = pgpages.RelationWriter()
    b = w.new_page(pgpages.PT_META)
    w.chain_write(pgpages.PT_META, b, struct.pack("<II", pgpages.TSV_VERSION, pgpages.TSV_MAGIC))     # (version, magic)
    path = str(tmp_path / "m")
    w.save(path)
    with pg.PgRelation(path) as rel:
        i = rel.info()
    assert (i["meta_magic"], i["meta_version"]) == (pgpages.TSV_MAGIC, pgpages.TSV_VERSION)


@pytest.mark.parametrize("order", [(0, 1, 2, 3), (2, 3, 0, 1)])
def test_plain_storage_relation_round_trips(pg, tmp_path, order):
    """storage_layout = plain: Node pages of PlainNode items (plain/node.rs:15-22) -> index_vectors + graph; the oracle's
    plain scan of the extracted snapshot equals its scan of the original."""
    from oracle import fixtures, oracle
    s = fixtures.to_plain(build_case(400, 48, L2, seed=6, R=12, L_build=24, deleted_every=8))
    path = str(tmp_path / "plain")
    meta, tids, _ = pgpages.write_index(s, path, field_order=order, dead_every=5)
    with pg.PgRelation(path) as rel:
        info = rel.info()
        assert info["pages_by_type"]["Node"] > 3 and "SbqNode" not in info["pages_by_type"]
        got, index_tid, _, layout = rel.extract_plain(pg.PgMeta(**meta))
        with pytest.raises(Exception) as e:
            rel.extract_sbq(pg.PgMeta(**meta))          # no SbqNode pages: nothing to extract is not an error, a bad start node is
        assert "start_nodes.default_node" in str(e.value)
    # the reader reports cells as (heap, main, neighbours, fourth); the writer's order is per DECLARED field
    assert layout == (order[3], order[0], order[2], order[1])
    assert np.array_equal(index_tid, tids) and got.storage_type == 1
    assert np.array_equal(got.index_vectors, s.index_vectors) and np.array_equal(got.nbrs, s.nbrs)
    assert np.array_equal(got.heap_tid, s.heap_tid) and got.start_default == s.start_default
    got.vectors = s.vectors
    q = fixtures.gen_vectors(5, 48, 2, "normal")
    a = oracle.scan_batch(s, q, None, None, 20, 0, 10)
    b = oracle.scan_batch(got, q, None, None, 20, 0, 10)
    assert np.array_equal(a[0], b[0])


def test_random_corruptions_never_crash_the_reader(pg, tmp_path):
    """300 random byte / word corruptions of page headers, line pointers and item bodies: every call returns a snapshot
    that passes its own invariants or a DANN_ERR_*; nothing reads outside the mapping (run under the ASan build of the
    emulated ABI as well: DANN_EMULATE=1 DANN_EMULATE_ASAN=1 LD_PRELOAD=libasan.so)."""
    from pgvectorscale_b200.diskann import DiskAnnError
    s = build_case(150, 48, L2, seed=9, R=12, L_build=24, labels=True)
    path = str(tmp_path / "rel")
    meta, _, w = pgpages.write_index(s, path)
    good = bytearray(open(path, "rb").read())
    rng = np.random.default_rng(123)
    nblocks = len(good) // 8192
    outcomes = {"ok": 0, "refused": 0}
    for it in range(300):
        blob = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            blk = int(rng.integers(0, nblocks))
            kind = int(rng.integers(0, 4))
            if kind == 0:                                   # page header field
                pos = blk * 8192 + int(rng.integers(0, 24))
            elif kind == 1:                                 # a line pointer
                pos = blk * 8192 + 24 + int(rng.integers(0, 64))
            elif kind == 2:                                 # the special area
                pos = blk * 8192 + 8184 + int(rng.integers(0, 8))
            else:                                           # anywhere in the item space (roots, relative pointers, lengths)
                pos = blk * 8192 + int(rng.integers(24, 8184))
            if rng.integers(0, 2):
                blob[pos] ^= 1 << int(rng.integers(0, 8))
            else:
                v = [0, 0xFF, 0x7F, 0x80][int(rng.integers(0, 4))]
                for j in range(int(rng.integers(1, 5))):
                    if pos + j < len(blob):
                        blob[pos + j] = v
        open(path, "wb").write(blob)
        try:
            with pg.PgRelation(path) as rel:
                rel.info()
                try:
                    rel.read_chain(0, 2)
                except DiskAnnError:
                    pass
                got, tids, _, _ = rel.extract_sbq(pg.PgMeta(**meta))
            valid = got.nbrs[got.nbrs != 0xFFFFFFFF]
            assert got.codes.shape == (got.n, s.words) and (valid < got.n).all() and len(tids) == got.n
            outcomes["ok"] += 1
        except DiskAnnError as e:
            assert e.code in (-7, -1, -5), e
            outcomes["refused"] += 1
    assert outcomes["refused"] > 50 and outcomes["ok"] > 20, outcomes


@pytest.mark.parametrize("dim", [16, 48, 500, 768])
def test_heap_vectors_inline_short_header_aligned_and_toasted(pg, tmp_path, dim):
    """The vector column of the table's relation files, by TID: 1-byte-header inline values (dim 16), 4-byte-header inline
    (48), out-of-line with a short last chunk (500: 1996 + 8 bytes) and the benchmark's 768 (two chunks); a bigint and a
    nullable text column in front; NULL vectors, vacuumed TIDs, a dead line pointer and a HOT redirect on the way."""
    import struct
    rng = np.random.default_rng(dim)
    n = 60
    x = rng.standard_normal((n, dim)).astype(np.float32)
    pre = []
    for i in range(n):
        tag = None if i % 5 == 0 else ("varlena", pgpages.varlena(b"t" * (i % 40)))
        pre.append([struct.pack("<q", 1000 + i), tag])
    heap_path, toast_path = str(tmp_path / "heap"), str(tmp_path / "toast")
    tids = pgpages.write_table(x, heap_path, toast_path, prefix_values=pre, prefix_atts=[(8, "d"), (-1, "i")],
                               null_rows={7, 31}, redirect_rows={11, 40})
    want = x.copy()
    want[[7, 31]] = 0
    ask = tids.copy()
    ask[3] &= np.uint64(0xFFFFFFFFFFFF0000)                 # vacuumed: offset 0
    want[3] = 0
    ask[20] = (ask[20] & np.uint64(0xFFFFFFFFFFFF0000)) | np.uint64(999)      # a line pointer that does not exist
    want[20] = 0
    with pg.PgRelation(heap_path) as heap, pg.PgRelation(toast_path) as toast:
        got, missing = pg.fetch_heap_vectors(heap, toast, ask, dim, [(8, "d"), (-1, "i")])
        assert missing == 4 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        if dim >= 500:
            with pytest.raises(Exception) as e:
                pg.fetch_heap_vectors(heap, None, ask, dim, [(8, "d"), (-1, "i")])
            assert "no TOAST relation" in str(e.value)
        with pytest.raises(Exception) as e:                  # the wrong dimension is refused, not reinterpreted
            pg.fetch_heap_vectors(heap, toast, ask, dim + 1, [(8, "d"), (-1, "i")])
        assert "not a vector(" in str(e.value)


def test_whole_loader_index_heap_toast_to_snapshot(pg, tmp_path):
    """index relation + table + TOAST files -> the complete snapshot a load takes; the oracle scans it like the original."""
    from oracle import fixtures, oracle
    s = build_case(300, 768, COSINE, seed=21, R=16, L_build=32)
    ipath, hpath, tpath = str(tmp_path / "idx"), str(tmp_path / "heap"), str(tmp_path / "toast")
    meta, _, _ = pgpages.write_index(s, ipath)
    htids = pgpages.write_table(s.vectors, hpath, tpath)
    s.heap_tid = htids                                       # the index points at where the rows really are
    meta, _, _ = pgpages.write_index(s, ipath)
    with pg.PgRelation(ipath) as rel, pg.PgRelation(hpath) as heap, pg.PgRelation(tpath) as toast:
        got, _, _, _ = rel.extract_sbq(pg.PgMeta(**meta))
        got.vectors, missing = pg.fetch_heap_vectors(heap, toast, got.heap_tid, 768)
    assert missing == 0 and np.array_equal(got.vectors, s.vectors)
    q = fixtures.gen_vectors(4, 768, 6, "normal")
    a = oracle.scan_batch(s, q, None, None, 30, 20, 10)
    b = oracle.scan_batch(got, q, None, None, 30, 20, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_relation_files_to_raw_snapshot_tool(pg, tmp_path, capsys):
    """tools/pg_to_snapshot.py: index + heap + TOAST files -> the DANNSNP1 file the sidecar and the C harnesses read."""
    import struct
    from pgvectorscale_b200.snapshot import Snapshot
    from tools import pg_to_snapshot
    s = build_case(250, 768, COSINE, seed=31, R=12, L_build=24)
    ipath, hpath, tpath, out = (str(tmp_path / x) for x in ("idx", "heap", "toast", "snap.raw"))
    pre = [[struct.pack("<q", i)] for i in range(s.n)]
    s.heap_tid = pgpages.write_table(s.vectors, hpath, tpath, prefix_values=pre, prefix_atts=[(8, "d")])
    meta, _, _ = pgpages.write_index(s, ipath)
    rc = pg_to_snapshot.main(["--index", ipath, "--heap", hpath, "--toast", tpath, "--out", out, "--dim", "768", "--R", str(s.R),
                              "--bits", str(s.bits), "--start", "%d:%d" % meta["start"], "--means", "%d:%d" % meta["means"], "--atts", "8d"])
    assert rc == 0 and "250 nodes, 0 heap rows gone" in capsys.readouterr().out
    blob = open(out, "rb").read()
    assert blob[:8] == Snapshot.RAW_MAGIC
    hdr = np.frombuffer(blob, dtype=np.uint64, count=16, offset=8)
    assert (int(hdr[0]), int(hdr[1]), int(hdr[5])) == (s.n, 768, s.R)
    assert s.vectors.tobytes() in blob and s.codes.tobytes() in blob
