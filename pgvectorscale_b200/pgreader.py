"""ctypes mirror of the relation-file reader (dann_pg_*, include/diskann_b200.h).

`PgRelation(path)` maps an index relation file, `.info()` is the page census + staleness fingerprint, `.read_chain()`
reassembles a chained item (util/chain.rs:125-183), `.extract_sbq(meta)` parses every SbqNode item into a `Snapshot`
(without heap vectors: those live in the table) plus the IndexPointer -> dense-id map.  Host only: nothing here touches
a GPU, and nothing here is a search path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import diskann
from .diskann import _check, _SnapshotDesc
from .snapshot import INVALID_NODE, Snapshot

TSV_MAGIC_NUMBER = 768756476   # meta_page.rs:22
PAGE_TYPES = ("MetaV1", "Node", "PqQuantizerDef", "PqQuantizerVector", "SbqMeansV1", "SbqNode", "MetaV2", "SbqMeans", "Meta")


class _RelationInfo(C.Structure):
    _fields_ = [("nblocks", C.c_uint32), ("pages_by_type", C.c_uint32 * 9), ("new_pages", C.c_uint32),
                ("foreign_pages", C.c_uint32), ("meta_magic", C.c_uint32), ("meta_version", C.c_uint32),
                ("node_items", C.c_uint64), ("max_lsn", C.c_uint64), ("fingerprint", C.c_uint64)]


class _PgMeta(C.Structure):
    _fields_ = [("num_dimensions", C.c_uint32), ("num_dimensions_to_index", C.c_uint32), ("bq_bits", C.c_uint32),
                ("num_neighbors", C.c_uint32), ("distance_type", C.c_int32), ("has_labels", C.c_int32),
                ("start_block", C.c_uint32), ("start_offset", C.c_uint16), ("n_start_labels", C.c_uint32),
                ("start_labels", C.c_void_p), ("start_label_block", C.c_void_p), ("start_label_offset", C.c_void_p),
                ("means_block", C.c_uint32), ("means_offset", C.c_uint16)]


class _HeapLayout(C.Structure):
    _fields_ = [("natts_before", C.c_uint32), ("attlen", C.c_void_p), ("attalign", C.c_char_p), ("dim", C.c_uint32),
                ("vector_align", C.c_char)]


class _PgSnapshot(C.Structure):
    _fields_ = [("snap", _SnapshotDesc), ("index_vectors", C.c_void_p), ("index_tid", C.c_void_p), ("fingerprint", C.c_uint64),
                ("layout", C.c_uint32 * 4), ("self", C.c_void_p)]


@dataclass
class PgMeta:
    """What MetaPage's getters return (meta_page.rs:212-282); pointers are (block, offset) pairs."""
    num_dimensions: int
    num_dimensions_to_index: int
    bq_bits: int
    num_neighbors: int
    distance_type: int
    has_labels: bool = False
    start: Optional[tuple] = None                      # start_nodes.default_node
    start_labels: dict = field(default_factory=dict)   # label -> (block, offset)
    means: Optional[tuple] = None                      # quantizer_metadata


def _bind(lib):
    if getattr(lib, "_pg_bound", False):
        return lib
    vp = C.c_void_p
    lib.dann_pg_relation_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.dann_pg_relation_close.argtypes = [vp]
    lib.dann_pg_relation_close.restype = None
    lib.dann_pg_relation_stat.argtypes = [vp, C.POINTER(_RelationInfo)]
    lib.dann_pg_read_chain.argtypes = [vp, C.c_uint32, C.c_uint16, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dann_pg_extract_sbq.argtypes = [vp, C.POINTER(_PgMeta), C.POINTER(C.POINTER(_PgSnapshot))]
    lib.dann_pg_extract_plain.argtypes = [vp, C.POINTER(_PgMeta), C.POINTER(C.POINTER(_PgSnapshot))]
    lib.dann_pg_snapshot_free.argtypes = [C.POINTER(_PgSnapshot)]
    lib.dann_pg_snapshot_free.restype = None
    lib.dann_pg_heap_fetch_vectors.argtypes = [vp, vp, C.POINTER(_HeapLayout), vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]
    lib._pg_bound = True
    return lib


def _copy(ptr, dtype, count):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr), dtype=dtype).copy()


class PgRelation:
    def __init__(self, path: str, lib_path: Optional[str] = None):
        self.lib = _bind(diskann.load_library(lib_path))
        self.h = C.c_void_p()
        _check(self.lib, self.lib.dann_pg_relation_open(path.encode(), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.dann_pg_relation_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def info(self) -> dict:
        o = _RelationInfo()
        _check(self.lib, self.lib.dann_pg_relation_stat(self.h, C.byref(o)))
        return {"nblocks": o.nblocks, "pages_by_type": {PAGE_TYPES[i]: int(o.pages_by_type[i]) for i in range(9) if o.pages_by_type[i]},
                "new_pages": o.new_pages, "foreign_pages": o.foreign_pages, "meta_magic": o.meta_magic,
                "meta_version": o.meta_version, "node_items": int(o.node_items), "max_lsn": int(o.max_lsn),
                "fingerprint": int(o.fingerprint)}

    def read_chain(self, block: int, offset: int, page_type: int = -1) -> bytes:
        n = C.c_size_t()
        _check(self.lib, self.lib.dann_pg_read_chain(self.h, block, offset, page_type, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(n.value, 1))
        _check(self.lib, self.lib.dann_pg_read_chain(self.h, block, offset, page_type, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def extract_sbq(self, meta: PgMeta):
        """memory_optimized layout -> (Snapshot with vectors=None, index_tid [n] uint64, fingerprint, layout tuple)"""
        return self._extract(meta, plain=False)

    def extract_plain(self, meta: PgMeta):
        """plain layout -> the same tuple; the Snapshot has storage_type=1 and index_vectors [n, dim_index]"""
        return self._extract(meta, plain=True)

    def _extract(self, meta: PgMeta, plain: bool):
        labs = sorted(meta.start_labels.items())
        sl = np.array([l for l, _ in labs], dtype=np.int16)
        sb = np.array([p[0] for _, p in labs], dtype=np.uint32)
        so = np.array([p[1] for _, p in labs], dtype=np.uint16)
        m = _PgMeta(meta.num_dimensions, meta.num_dimensions_to_index, meta.bq_bits, meta.num_neighbors, meta.distance_type,
                    int(meta.has_labels), meta.start[0] if meta.start else INVALID_NODE, meta.start[1] if meta.start else 0,
                    len(labs), sl.ctypes.data if len(labs) else None, sb.ctypes.data if len(labs) else None,
                    so.ctypes.data if len(labs) else None, meta.means[0] if meta.means else INVALID_NODE,
                    meta.means[1] if meta.means else 0)
        out = C.POINTER(_PgSnapshot)()
        fn = self.lib.dann_pg_extract_plain if plain else self.lib.dann_pg_extract_sbq
        _check(self.lib, fn(self.h, C.byref(m), C.byref(out)))
        try:
            d = out.contents.snap
            n, R, words = d.n, d.R, d.words
            nl = _copy(d.label_off, np.uint32, n + 1) if d.has_labels else None
            snap = Snapshot(
                n=n, dim=d.dim, dim_index=d.dim_index, bits=d.bits, words=words, R=R, distance_type=d.distance_type,
                has_labels=bool(d.has_labels), count=int(d.count),
                mean=_copy(d.mean, np.float32, d.dim_index) if d.mean else np.zeros(d.dim_index, np.float32),
                m2=_copy(d.m2, np.float32, d.dim_index) if d.m2 else None,
                codes=_copy(d.codes, np.uint64, n * words).reshape(n, words),
                nbrs=_copy(d.nbrs, np.uint32, n * R).reshape(n, R), heap_tid=_copy(d.heap_tid, np.uint64, n), vectors=None,
                start_default=d.start_default,
                start_labels=_copy(d.start_labels, np.int16, d.n_start_labels) if d.n_start_labels else None,
                start_label_nodes=_copy(d.start_label_nodes, np.uint32, d.n_start_labels) if d.n_start_labels else None,
                label_off=nl, labels=_copy(d.labels, np.int16, int(nl[-1])) if nl is not None else None)
            if plain:
                snap.storage_type = 1
                snap.index_vectors = _copy(out.contents.index_vectors, np.float32, n * d.dim_index).reshape(n, d.dim_index)
            index_tid = _copy(out.contents.index_tid, np.uint64, n)
            return snap, index_tid, int(out.contents.fingerprint), tuple(out.contents.layout)
        finally:
            self.lib.dann_pg_snapshot_free(out)


def fetch_heap_vectors(heap: PgRelation, toast: Optional[PgRelation], heap_tid: np.ndarray, dim: int,
                       atts_before: Sequence[tuple] = ()):
    """Vector column of the table's relation file(s) for the rows heap_tid[i] -> (vectors [n, dim] f32, rows missing).
    atts_before: (attlen, attalign) of every column in front of the vector column, e.g. [(8, 'd')] for a bigint id."""
    tids = np.ascontiguousarray(heap_tid, dtype=np.uint64)
    al = np.array([a for a, _ in atts_before], dtype=np.int16)
    aa = "".join(c for _, c in atts_before).encode()
    lay = _HeapLayout(len(atts_before), al.ctypes.data if len(al) else None, aa if aa else None, dim, b"\0")
    out = np.zeros((len(tids), dim), np.float32)
    miss = C.c_uint32()
    _check(heap.lib, heap.lib.dann_pg_heap_fetch_vectors(heap.h, toast.h if toast else None, C.byref(lay), tids.ctypes.data,
                                                         len(tids), out.ctypes.data, C.byref(miss)))
    return out, int(miss.value)
