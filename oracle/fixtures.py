"""Test fixtures built THROUGH THE ORACLE (test infrastructure only).

make_index() restates what `CREATE INDEX ... USING diskann` leaves behind for the scan path:
train SbqMeans over the heap-scan order, quantize every vector, serial Vamana build over the
codes, and package the result as a pgvectorscale_b200.snapshot.Snapshot.
"""
from __future__ import annotations

import numpy as np

from pgvectorscale_b200.snapshot import (COSINE, INVALID_NODE, Snapshot, code_words,
                                         default_bits, make_heap_tids)

from . import oracle


def gen_vectors(n, dim, seed, kind="uniform"):
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "uniform":          # mirrors the reference's random() fixtures (build.rs:1226)
        return rng.random((n, dim), dtype=np.float32)
    v = rng.standard_normal((n, dim), dtype=np.float32)   # "Cohere-shape": unit norm
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def gen_labels(n, seed, max_label=16):
    """1-2 labels uniform in 1..max_label per node (mirrors build.rs:1988-1991)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    off = np.zeros(n + 1, np.uint32)
    out = []
    for i in range(n):
        k = int(rng.integers(1, 3))
        ls = sorted(set(int(x) for x in rng.integers(1, max_label + 1, size=k)))
        out.extend(ls)
        off[i + 1] = len(out)
    return off, np.asarray(out, np.int16)


def make_index(vectors, distance_type=COSINE, bits=None, R=50, L_build=100, alpha=1.2,
               dim_index=None, label_off=None, labels=None, train_on_data=True):
    vectors = np.ascontiguousarray(vectors, np.float32)
    n, dim = vectors.shape
    dim_index = dim if dim_index is None else dim_index
    bits = default_bits(dim_index) if bits is None else bits
    words = code_words(dim_index, bits)
    # what the build sees: truncated to dim_index, then cosine-normalised (pg_vector.rs:143-155)
    idx = vectors[:, :dim_index].copy()
    if distance_type == COSINE:
        for i in range(n):
            idx[i] = oracle.preprocess_cosine(idx[i])
    if train_on_data and n > 0:
        mean, m2, count = oracle.train(idx, bits)
    else:  # index created on an empty table: count=0, zeros (build.rs:1419-1433)
        mean, m2, count = np.zeros(dim_index, np.float32), np.zeros(dim_index, np.float32), 0
    codes = oracle.quantize_all(idx, bits, mean, m2, count) if n else np.zeros((0, words), np.uint64)
    if n:
        nbrs, sd, sl, sn = oracle.build_graph(codes, R, L_build, alpha, label_off, labels)
    else:
        nbrs, sd, sl, sn = np.zeros((0, R), np.uint32), INVALID_NODE, None, None
    return Snapshot(n=n, dim=dim, dim_index=dim_index, bits=bits, words=words, R=R,
                    distance_type=distance_type, has_labels=label_off is not None, count=count,
                    mean=mean, m2=m2, codes=codes, nbrs=nbrs, heap_tid=make_heap_tids(n),
                    vectors=vectors, start_default=sd,
                    start_labels=sl if label_off is not None else None,
                    start_label_nodes=sn if label_off is not None else None,
                    label_off=label_off, labels=labels)


def to_plain(s):
    """Re-express an SBQ fixture as a plain-storage index over the same graph: each node keeps the f32 vector
    PlainStorage would have stored (plain/node.rs:17-22 via pg_vector.rs:143-148: truncated to dim_index, and
    cosine-normalised after truncation).  The reference rejects inner product and labels for this layout
    (build.rs:264-290); the graph itself would be built on exact distances there - any valid graph serves
    scan parity."""
    import copy
    from . import oracle
    p = copy.copy(s)
    iv = np.ascontiguousarray(s.vectors[:, :s.dim_index], dtype=np.float32).copy()
    if s.distance_type == COSINE:
        for i in range(len(iv)):
            iv[i] = oracle.preprocess_cosine(iv[i])
    p.storage_type = 1
    p.index_vectors = iv
    p.has_labels = 0
    p.label_off = p.labels = None
    p.start_labels = p.start_label_nodes = None
    return p
