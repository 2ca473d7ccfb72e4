"""Flat index snapshot: the SBQ storage layout kept at the drop-in boundary.

One snapshot holds everything the index-scan path reads from a pgvectorscale `diskann`
index (reference paths relative to /root/reference/pgvectorscale/src/access_method/):

  codes[n, words]   u64  SBQ code of each node; bit p = i*bits+j lives in word p//64,
                         bit p%64 (LSB first)                      sbq/quantize.rs:57-62,73-86
  nbrs[n, R]        u32  dense neighbour ids, list ends at the first 0xFFFFFFFF
                         (InvalidBlockNumber sentinel)             sbq/node.rs:261-285,389-394
  heap_tid[n]       u64  (block << 16) | offset ; offset == 0 marks a vacuumed node
                         sbq/node.rs:149-152, scan.rs:231-234
  vectors[n, dim]   f32  heap column values (raw; the loader cosine-normalises once with
                         the reference's arithmetic)               sbq/storage.rs:304-328
  mean/m2/count          SbqMeans                                  sbq/mod.rs:77-121
  start_default, start_labels -> start_label_nodes                 graph/start_nodes.rs:15-48
  label_off[n+1], labels  CSR of each node's sorted label set      labels/mod.rs:17-37
  dim, dim_index, bits, R, distance_type, has_labels               meta_page.rs:181-282
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Optional

import numpy as np

INVALID_NODE = 0xFFFFFFFF
COSINE, L2, IP = 0, 1, 2  # distance/mod.rs:11-15


def code_words(dim_index: int, bits: int) -> int:
    """sbq/quantize.rs:38-46 quantized_size_internal."""
    nb = dim_index * bits
    return nb // 64 if nb % 64 == 0 else nb // 64 + 1


def default_bits(dim_index: int) -> int:
    """meta_page.rs:312-323: 2 bits/dim below 900 dimensions, else 1."""
    return 2 if dim_index < 900 else 1


@dataclass
class Snapshot:
    n: int
    dim: int
    dim_index: int
    bits: int
    words: int
    R: int
    distance_type: int
    has_labels: bool
    count: int
    mean: np.ndarray
    m2: Optional[np.ndarray]
    codes: np.ndarray
    nbrs: np.ndarray
    heap_tid: np.ndarray
    vectors: Optional[np.ndarray]   # None: supplied after load (DiskAnnIndex.set_vectors)
    start_default: int = INVALID_NODE
    start_labels: Optional[np.ndarray] = None
    start_label_nodes: Optional[np.ndarray] = None
    label_off: Optional[np.ndarray] = None
    labels: Optional[np.ndarray] = None
    # storage_layout (storage.rs:144-169): 0 = SBQ ("memory_optimized", the default and the only layout the CUDA
    # path scans today), 1 = plain: nodes hold index_vectors[n, dim_index] f32 (plain/node.rs:17-22).  The
    # oracle restates both; the plain scan kernel is the next scope row.
    storage_type: int = 0
    index_vectors: Optional[np.ndarray] = None

    def validate(self) -> None:
        assert self.words == code_words(self.dim_index, self.bits)
        assert self.codes.shape == (self.n, self.words) and self.codes.dtype == np.uint64
        assert self.nbrs.shape == (self.n, self.R) and self.nbrs.dtype == np.uint32
        assert self.heap_tid.shape == (self.n,) and self.heap_tid.dtype == np.uint64
        if self.vectors is not None:
            assert self.vectors.shape == (self.n, self.dim) and self.vectors.dtype == np.float32
        assert self.mean.shape == (self.dim_index,)
        assert 1 <= self.dim_index <= self.dim
        if self.has_labels:
            assert self.label_off is not None and self.label_off.shape == (self.n + 1,)

    def save(self, path: str) -> None:
        d = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                d[f.name] = v
        np.savez(path, **d)

    RAW_MAGIC = b"DANNSNP1"
    RAW_ARRAYS = (("mean", np.float32), ("m2", np.float32), ("codes", np.uint64), ("nbrs", np.uint32),
                  ("heap_tid", np.uint64), ("vectors", np.float32), ("start_labels", np.int16),
                  ("start_label_nodes", np.uint32), ("label_off", np.uint32), ("labels", np.int16),
                  ("index_vectors", np.float32))

    def save_raw(self, path: str) -> None:
        """Flat binary form for C hosts (harness/snapshot_raw.h reads it into a dann_snapshot_desc):
        8-byte magic, 16 little-endian u64 header words, then every array of RAW_ARRAYS in order, each preceded by its
        byte length (u64) and padded to a multiple of 64 bytes (absent arrays have length 0)."""
        hdr = np.zeros(16, np.uint64)
        hdr[:12] = [self.n, self.dim, self.dim_index, self.bits, self.words, self.R, int(self.distance_type) & 0xFFFFFFFF,
                    int(bool(self.has_labels)), int(self.count), int(self.start_default),
                    0 if self.start_labels is None else len(self.start_labels), int(self.storage_type or 0)]
        with open(path, "wb") as f:
            f.write(self.RAW_MAGIC)
            f.write(hdr.tobytes())
            for name, dt in self.RAW_ARRAYS:
                v = getattr(self, name)
                b = b"" if v is None else np.ascontiguousarray(v, dtype=dt).tobytes()
                f.write(np.uint64(len(b)).tobytes())
                f.write(b)
                f.write(b"\0" * ((-len(b)) % 64))

    @classmethod
    def load(cls, path: str) -> "Snapshot":
        z = np.load(path)
        kw = {}
        for f in fields(cls):
            if f.name in z.files:
                v = z[f.name]
                kw[f.name] = v.item() if v.ndim == 0 else v
            else:
                kw[f.name] = None
        kw["has_labels"] = bool(kw["has_labels"])
        return cls(**kw)


def make_heap_tids(n: int, tuples_per_page: int = 2) -> np.ndarray:
    """Synthetic heap TIDs (block, offset>=1) laid out like a heap scan of 3 KB vectors."""
    i = np.arange(n, dtype=np.uint64)
    return ((i // tuples_per_page) << np.uint64(16)) | (i % tuples_per_page + np.uint64(1))
