"""Test infrastructure: writes the relation file of a pgvectorscale `diskann` index from a Snapshot, following the
reference's WRITE path rule by rule (paths relative to /root/reference/pgvectorscale/src/):

  * PageInit(page, BLCKSZ, 4) + TsvPageOpaqueData in the special area               util/page.rs:59-69,127-136
  * PageAddItemExtended: line pointer (lp_off:15, lp_flags:2, lp_len:15), item at MAXALIGN'ed pd_upper - size
  * ChainTapeWriter::write: 8-byte ArchivedChainItemHeader in front of every chunk, a chunk takes all the aligned free
    space of its page, `next` = (next block, 1), the last chunk ends with an invalid pointer     util/chain.rs:76-122
  * Tape::write: a node item never splits; a page that cannot take it is followed by a new one   util/tape.rs:50-72
  * MetaPage::store: block 0, item 1 = MetaPageHeader, item 2 = MetaPage                        meta_page.rs:360-384
  * rkyv 0.7 archives: out-of-line data of the fields in declaration order, the root struct last; ArchivedVec =
    { i32 offset relative to the field, u32 len }, ArchivedItemPointer = { u32, u16, pad }       sbq/node.rs:26-42

Nothing here is product code; the reader under test is pgvectorscale_b200/csrc/dann_pgreader.h.  `field_order`
permutes the cells of the archived node root to stand in for a rustc that reorders repr(Rust) fields.
"""
from __future__ import annotations

import struct

import numpy as np

BLCKSZ = 8192
HDR = 24
SPECIAL = BLCKSZ - 8
TSV_PAGE_ID = 0xAE24
TSV_MAGIC = 768756476
TSV_VERSION = 3
INVALID_BLOCK = 0xFFFFFFFF
PT_NODE, PT_SBQ_NODE, PT_SBQ_MEANS, PT_META = 1, 5, 7, 8


def _align8(x):
    return (x + 7) & ~7


class Page:
    def __init__(self, page_type: int, lsn: int = 0):
        self.b = bytearray(BLCKSZ)
        self.lower, self.upper = HDR, SPECIAL
        self.b[SPECIAL] = page_type
        struct.pack_into("<H", self.b, SPECIAL + 2, TSV_PAGE_ID)
        self.lsn = lsn
        self.nitems = 0

    def aligned_free(self) -> int:
        """PageGetFreeSpace (room for one more line pointer), rounded down to 8: get_aligned_free_space."""
        free = self.upper - self.lower - 4
        free = max(free, 0)
        return free - free % 8

    def add_item(self, data: bytes, flags: int = 1) -> int:
        size = len(data)
        up = self.upper - _align8(size)
        assert up >= self.lower + 4, "item does not fit"
        self.b[up:up + size] = data
        struct.pack_into("<I", self.b, self.lower, up | (flags << 15) | (size << 17))
        self.lower += 4
        self.upper = up
        self.nitems += 1
        return self.nitems

    def bytes(self) -> bytes:
        struct.pack_into("<IIHHHHHHI", self.b, 0, self.lsn >> 32, self.lsn & 0xFFFFFFFF, 0, 0, self.lower, self.upper, SPECIAL,
                         BLCKSZ | 4, 0)
        return bytes(self.b)


def item_pointer(block: int, offset: int) -> bytes:
    return struct.pack("<IHH", block, offset, 0)


class RelationWriter:
    def __init__(self):
        self.pages = []
        self.lsn = 0x1000

    def new_page(self, page_type: int) -> int:
        self.lsn += 0x28
        self.pages.append(Page(page_type, self.lsn))
        return len(self.pages) - 1

    def chain_write(self, page_type: int, start_block: int, data: bytes):
        """ChainTapeWriter::write -> (block, offset) of the first chunk"""
        cur = start_block
        pg = self.pages[cur]
        if pg.aligned_free() < 8 + 1:
            cur = self.new_page(page_type)
            pg = self.pages[cur]
        first = None
        while 8 + len(data) > pg.aligned_free():
            nxt = self.new_page(page_type)
            size = pg.aligned_free() - 8
            off = pg.add_item(item_pointer(nxt, 1) + data[:size])
            first = first or (cur, off)
            data = data[size:]
            cur, pg = nxt, self.pages[nxt]
        off = pg.add_item(item_pointer(INVALID_BLOCK, 0) + data)
        return first or (cur, off), cur

    def tape_write(self, page_type: int, cur: int, data: bytes):
        pg = self.pages[cur]
        if pg.aligned_free() < len(data):
            cur = self.new_page(page_type)
            pg = self.pages[cur]
        return (cur, pg.add_item(data)), cur

    def save(self, path: str, segment_blocks: int = 131072):
        blob = b"".join(p.bytes() for p in self.pages)
        seg = segment_blocks * BLCKSZ
        for i in range(0, max(len(blob), 1), seg):
            with open(path if i == 0 else f"{path}.{i // seg}", "wb") as f:
                f.write(blob[i:i + seg])


def archive_node(heap_tid: int, code: np.ndarray, nbr_ptrs, fourth: bytes, n_fourth: int, field_order=(0, 1, 2, 3)) -> bytes:
    """rkyv archive of Classic/LabeledSbqNode: [code u64s][neighbour ItemPointers][fourth's elements][32-byte root].
    field_order[i] = which cell of the root field i (heap pointer, code, neighbours, fourth) lands in."""
    body = bytearray()
    code_pos = len(body)
    body += np.ascontiguousarray(code, dtype=np.uint64).tobytes()
    nbr_pos = len(body)
    for b, o in nbr_ptrs:
        body += item_pointer(b, o)
    fourth_pos = len(body)
    body += fourth
    while len(body) % 4:
        body += b"\0"
    root = len(body)
    cells = [None] * 4

    def vec(pos, n, cell):
        return struct.pack("<iI", pos - (root + 8 * cell), n)

    cells[field_order[0]] = item_pointer(heap_tid >> 16, heap_tid & 0xFFFF)
    cells[field_order[1]] = vec(code_pos, len(code), field_order[1])
    cells[field_order[2]] = vec(nbr_pos, len(nbr_ptrs), field_order[2])
    cells[field_order[3]] = vec(fourth_pos, n_fourth, field_order[3])
    return bytes(body) + b"".join(cells)


def archive_plain_node(heap_tid: int, vector: np.ndarray, nbr_ptrs, field_order=(0, 1, 2, 3)) -> bytes:
    """rkyv archive of PlainNode { vector: Vec<f32>, pq_vector: Vec<u8> (empty), neighbor_index_pointers, heap_item_pointer }
    (plain/node.rs:15-22): [f32s][neighbour ItemPointers][32-byte root].  field_order[i] = the root cell of declared field i."""
    body = bytearray()
    body += np.ascontiguousarray(vector, dtype=np.float32).tobytes()
    nbr_pos = len(body)
    for b, o in nbr_ptrs:
        body += item_pointer(b, o)
    pq_pos = len(body)
    while len(body) % 4:
        body += b"\0"
    root = len(body)
    cells = [None] * 4

    def vec(pos, n, cell):
        return struct.pack("<iI", pos - (root + 8 * cell), n)

    cells[field_order[0]] = vec(0, len(vector), field_order[0])
    cells[field_order[1]] = vec(pq_pos, 0, field_order[1])
    cells[field_order[2]] = vec(nbr_pos, len(nbr_ptrs), field_order[2])
    cells[field_order[3]] = item_pointer(heap_tid >> 16, heap_tid & 0xFFFF)
    return bytes(body) + b"".join(cells)


def archive_means(count: int, mean: np.ndarray, m2) -> bytes:
    body = bytearray()
    mpos = 0
    body += np.ascontiguousarray(mean, dtype=np.float32).tobytes()
    m2pos = len(body)
    nm2 = 0
    if m2 is not None:
        body += np.ascontiguousarray(m2, dtype=np.float32).tobytes()
        nm2 = len(m2)
    while len(body) % 8:
        body += b"\0"
    root = len(body)
    return bytes(body) + struct.pack("<Q", count) + struct.pack("<iI", mpos - (root + 8), len(mean)) + \
        struct.pack("<iI", m2pos - (root + 16), nm2)


def write_index(snap, path: str, field_order=(0, 1, 2, 3), meta_body: bytes = b"\0" * 96, dead_every: int = 0, segment_blocks: int = 131072):
    """Snapshot -> relation file(s).  Returns (PgMeta kwargs, index_tid[n]) - what the host side would know from MetaPage.
    dead_every: every such item slot on node pages is an LP_DEAD line pointer (not a node), to exercise the id map."""
    w = RelationWriter()
    meta_block = w.new_page(PT_META)
    ip, _ = w.chain_write(PT_META, meta_block, struct.pack("<II", TSV_MAGIC, TSV_VERSION))
    assert ip == (0, 1)
    ip, _ = w.chain_write(PT_META, meta_block, meta_body)      # the MetaPage archive: opaque to the reader
    assert ip == (0, 2)
    # node items are written with invalid neighbours first, their IndexPointers collected, then the neighbour pointers are
    # patched in place - the order a build does it in (Tape::write, then the neighbour lists are rewritten)
    n, R = snap.n, snap.R
    plain = getattr(snap, "storage_type", 0) == 1
    node_pt = PT_NODE if plain else PT_SBQ_NODE
    cur = w.new_page(node_pt)
    tids, where = [], []
    slot = 0
    for i in range(n):
        slot += 1
        if dead_every and slot % dead_every == 0 and w.pages[cur].aligned_free() >= 8:
            w.pages[cur].add_item(b"\0" * 8, flags=3)       # LP_DEAD: not a node, takes an offset number
        if snap.has_labels:
            ls = snap.labels[snap.label_off[i]:snap.label_off[i + 1]]
            fourth, nf = np.ascontiguousarray(ls, dtype=np.int16).tobytes(), len(ls)
        else:
            fourth, nf = b"", 0
        if plain:
            data = archive_plain_node(int(snap.heap_tid[i]), snap.index_vectors[i], [(INVALID_BLOCK, 0)] * R, field_order)
            nbr_at = snap.dim_index * 4
        else:
            data = archive_node(int(snap.heap_tid[i]), snap.codes[i], [(INVALID_BLOCK, 0)] * R, fourth, nf, field_order)
            nbr_at = snap.words * 8
        ip, cur = w.tape_write(node_pt, cur, data)
        tids.append(ip)
        where.append((cur, w.pages[cur].upper + nbr_at))
    for i in range(n):
        blk, pos = where[i]
        for j in range(R):
            v = int(snap.nbrs[i, j])
            if v != 0xFFFFFFFF:
                w.pages[blk].b[pos + 8 * j:pos + 8 * j + 8] = item_pointer(*tids[v])
    means = None
    if snap.mean is not None and not plain:
        mb = w.new_page(PT_SBQ_MEANS)
        means, _ = w.chain_write(PT_SBQ_MEANS, mb, archive_means(int(snap.count), snap.mean, snap.m2))
    w.save(path, segment_blocks)
    meta = dict(num_dimensions=snap.dim, num_dimensions_to_index=snap.dim_index, bq_bits=snap.bits, num_neighbors=R,
                distance_type=int(snap.distance_type), has_labels=bool(snap.has_labels),
                start=None if snap.start_default == 0xFFFFFFFF else tids[int(snap.start_default)],
                start_labels={} if snap.start_labels is None else
                {int(l): tids[int(v)] for l, v in zip(snap.start_labels, snap.start_label_nodes)},
                means=means)
    return meta, np.array([(b << 16) | o for b, o in tids], dtype=np.uint64), w


# ---- heap and TOAST relations (Postgres' own formats: htup_details.h, postgres.h varlena, detoast.h, heaptoast.h) ----------
TOAST_MAX_CHUNK_SIZE = 1996


class HeapPage(Page):
    """PageInit(page, BLCKSZ, 0): no special space."""

    def __init__(self, lsn: int = 0):
        self.b = bytearray(BLCKSZ)
        self.lower, self.upper = HDR, BLCKSZ
        self.lsn = lsn
        self.nitems = 0

    def bytes(self) -> bytes:
        struct.pack_into("<IIHHHHHHI", self.b, 0, self.lsn >> 32, self.lsn & 0xFFFFFFFF, 0, 0, self.lower, self.upper, BLCKSZ,
                         BLCKSZ | 4, 0)
        return bytes(self.b)

    def add_redirect(self, to_offset: int) -> int:
        struct.pack_into("<I", self.b, self.lower, to_offset | (2 << 15))      # LP_REDIRECT: lp_off = target offset number
        self.lower += 4
        self.nitems += 1
        return self.nitems


def _align(off, a):
    n = {"c": 1, "s": 2, "i": 4, "d": 8}[a]
    return (off + n - 1) & ~(n - 1)


def varlena(payload: bytes, allow_short=True) -> bytes:
    """what heap_fill_tuple stores for an inline value: 1-byte header when the whole thing fits 127 bytes, else 4-byte"""
    if allow_short and len(payload) + 1 <= 0x7F:
        return bytes([((len(payload) + 1) << 1) | 1]) + payload
    return struct.pack("<I", (len(payload) + 4) << 2) + payload


def toast_pointer(rawsize: int, extsize: int, valueid: int, toastrelid: int) -> bytes:
    return bytes([0x01, 18]) + struct.pack("<iIII", rawsize, extsize, valueid, toastrelid)


def heap_tuple(values, atts, block: int, offset: int) -> bytes:
    """values[i]: None (NULL), bytes for a fixed-width attribute, or ("varlena", stored bytes incl. header) / ("pointer", bytes).
    atts[i] = (attlen, attalign).  heap_fill_tuple's layout rules."""
    natts = len(values)
    hasnull = any(v is None for v in values)
    hoff = 23 + ((natts + 7) // 8 if hasnull else 0)
    hoff = (hoff + 7) & ~7
    body = bytearray()
    infomask = 0x0100 | 0x0800                                 # XMIN_COMMITTED | XMAX_INVALID
    for v, (attlen, attalign) in zip(values, atts):
        if v is None:
            continue
        off = hoff + len(body)
        if attlen > 0:
            body += b"\0" * (_align(off, attalign) - off) + v
        else:
            kind, data = v
            infomask |= 0x0002                                  # HASVARWIDTH
            if kind == "pointer":
                infomask |= 0x0004                              # HASEXTERNAL
            short = kind == "pointer" or (data[0] & 1)
            if not short:
                body += b"\0" * (_align(off, attalign) - off)   # 4-byte headers are aligned, 1-byte ones are not
            body += data
    if hasnull:
        infomask |= 0x0001
    hdr = bytearray(hoff)
    struct.pack_into("<IIIHHHHHB", hdr, 0, 700, 0, 0, block >> 16, block & 0xFFFF, offset, natts, infomask, hoff)
    if hasnull:
        for i, v in enumerate(values):
            if v is not None:
                hdr[23 + (i >> 3)] |= 1 << (i & 7)
    return bytes(hdr) + bytes(body)


class HeapWriter:
    def __init__(self):
        self.pages = [HeapPage(0x2000)]

    def add(self, make_tuple) -> tuple:
        """make_tuple(block, offset) -> bytes; returns the TID"""
        pg = self.pages[-1]
        probe = make_tuple(len(self.pages) - 1, pg.nitems + 1)
        if pg.upper - pg.lower - 4 < _align8(len(probe)):
            self.pages.append(HeapPage(0x2000 + 0x30 * len(self.pages)))
            pg = self.pages[-1]
        blk = len(self.pages) - 1
        return (blk, pg.add_item(make_tuple(blk, pg.nitems + 1)))

    def save(self, path):
        with open(path, "wb") as f:
            for p in self.pages:
                f.write(p.bytes())


def write_table(vectors, path_heap, path_toast, prefix_values=None, prefix_atts=(), toastrelid=16999, null_rows=(), redirect_rows=(),
                force_external_above=2000):
    """vectors [n, dim] -> heap (+ TOAST) relation files of a table (prefix columns..., embedding vector(dim)).
    Values above force_external_above bytes go out of line (TOAST_TUPLE_THRESHOLD); returns heap_tid [n] uint64.
    redirect_rows: rows whose TID is an LP_REDIRECT to the real tuple (a HOT chain's root after pruning)."""
    heap, toast = HeapWriter(), HeapWriter()
    toast_atts = [(4, "i"), (4, "i"), (-1, "i")]
    tids = []
    next_value = 24000
    for i, x in enumerate(vectors):
        payload = struct.pack("<hh", len(x), 0) + np.ascontiguousarray(x, dtype=np.float32).tobytes()
        pre = list(prefix_values[i]) if prefix_values is not None else []
        if i in null_rows:
            val = None
        elif len(payload) + 4 > force_external_above:
            vid = next_value
            next_value += 1
            for seq, s in enumerate(range(0, len(payload), TOAST_MAX_CHUNK_SIZE)):
                chunk = payload[s:s + TOAST_MAX_CHUNK_SIZE]
                toast.add(lambda b, o, seq=seq, chunk=chunk: heap_tuple(
                    [struct.pack("<I", vid), struct.pack("<i", seq), ("varlena", varlena(chunk))], toast_atts, b, o))
            val = ("pointer", toast_pointer(len(payload) + 4, len(payload), vid, toastrelid))
        else:
            val = ("varlena", varlena(payload))
        atts = list(prefix_atts) + [(-1, "i")]
        if i in redirect_rows:
            pg = heap.pages[-1]
            if pg.upper - pg.lower < 600 + len(payload):
                heap.pages.append(HeapPage(0x2000 + 0x30 * len(heap.pages)))
                pg = heap.pages[-1]
            blk = len(heap.pages) - 1
            root = pg.add_redirect(pg.nitems + 2)
            real = pg.add_item(heap_tuple(pre + [val], atts, blk, root + 1))
            assert real == root + 1
            tids.append((blk, root))
        else:
            tids.append(heap.add(lambda b, o: heap_tuple(pre + [val], atts, b, o)))
    heap.save(path_heap)
    toast.save(path_toast)
    return np.array([(b << 16) | o for b, o in tids], dtype=np.uint64)
