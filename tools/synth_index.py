"""Synthetic diskann index fixtures at benchmark scale (1M+ nodes), built on the GPU with torch.

NOT the product and NOT the reference's build: index construction is out of the hot-path
scope.  The serial restatement of the reference build
(oracle.build_graph) needs hours at 1M nodes, so bench.py gets its graph from this batch
builder instead: exact kNN candidates by chunked matmul, a Vamana-style alpha prune
(restating the *shape* of graph/mod.rs:392-488 with a single alpha), reverse edges, second
prune.  Parity does not depend on how the graph was made: the oracle and the CUDA path search
the SAME snapshot; graph quality only moves recall.

SBQ codes of the nodes are produced by the product's own quantizer kernel
(dann_prepare_queries: the reference quantizes nodes and queries with the same function,
sbq/quantize.rs:52-102).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from pgvectorscale_b200.snapshot import COSINE, INVALID_NODE, Snapshot, code_words, default_bits, make_heap_tids


def gen_dataset(n: int, dim: int, seed: int, kind: str = "lowrank", device="cuda", rank: int = 32,
                eta: float = 0.5, basis_seed: int = 0x5EED00AA) -> torch.Tensor:
    """Unit-norm f32 vectors.
    kind="gaussian": N(0,1) per dimension, L2-normalised (SURVEY §8d config 2 as written).
    kind="lowrank" : "Cohere-shape": a rank-`rank` Gaussian latent mapped to `dim` dimensions by a
                     fixed random basis plus isotropic noise of relative size `eta`, normalised.
                     Real text embeddings have a low intrinsic dimension and nearest-neighbour
                     cosine ~0.6-0.9; i.i.d. Gaussians in 768-d have nearest neighbours at cosine
                     ~0.17, indistinguishable under ANY 1-2 bit quantizer within rescore<=1000."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if kind == "gaussian":
        x = torch.randn((n, dim), generator=g, device=device, dtype=torch.float32)
    elif kind == "lowrank":
        gb = torch.Generator(device=device)
        gb.manual_seed(basis_seed)
        basis = torch.randn((rank, dim), generator=gb, device=device, dtype=torch.float32)
        z = torch.randn((n, rank), generator=g, device=device, dtype=torch.float32)
        u = z @ basis
        u = u / u.norm(dim=1, keepdim=True)
        noise = torch.randn((n, dim), generator=g, device=device, dtype=torch.float32)
        noise = noise / noise.norm(dim=1, keepdim=True)
        x = u + eta * noise
    elif kind == "uniform":
        x = torch.rand((n, dim), generator=g, device=device, dtype=torch.float32)
        return x
    else:
        raise ValueError(kind)
    return x / x.norm(dim=1, keepdim=True)


@torch.no_grad()
def knn_candidates(x: torch.Tensor, k: int, chunk: int = 2048):
    """Exact top-k by cosine (bf16 tensor-core matmul; candidates only). -> idx [n,k] int64, sim [n,k] f32"""
    n = x.shape[0]
    xb = x.to(torch.bfloat16)
    idx = torch.empty((n, k), dtype=torch.int64, device=x.device)
    sim = torch.empty((n, k), dtype=torch.float32, device=x.device)
    kq = max(min(k, n - 1), 0)
    idx.fill_(-1)
    sim.fill_(-4.0)
    if kq == 0:
        return idx, sim
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        sc = (xb[s:e] @ xb.T).float()
        sc[torch.arange(e - s, device=x.device), torch.arange(s, e, device=x.device)] = -4.0   # drop self
        v, i = torch.topk(sc, kq, dim=1)
        idx[s:e, :kq] = i
        sim[s:e, :kq] = v
    return idx, sim


@torch.no_grad()
def alpha_prune(x: torch.Tensor, cand: torch.Tensor, R: int, alpha: float = 1.2, chunk: int = 2048):
    """cand [n,C] int64 (-1 = empty), any order. Keeps <= R per node: walk candidates by increasing
    distance, keep one unless an already kept neighbour s has alpha*d(s,c) <= d(p,c)."""
    n, C = cand.shape
    xb = x.to(torch.bfloat16)
    out = torch.full((n, R), -1, dtype=torch.int64, device=x.device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        c = cand[s:e]
        valid = c >= 0
        cs = c.clamp(min=0)
        xc = xb[cs]                                              # [m,C,d]
        dpc = 1.0 - torch.einsum("mcd,md->mc", xc, xb[s:e]).float()
        dpc = torch.where(valid, dpc, torch.full_like(dpc, 1e9))
        order = torch.argsort(dpc, dim=1, stable=True)
        c = torch.gather(c, 1, order)
        dpc = torch.gather(dpc, 1, order)
        xc = torch.gather(xc, 1, order[:, :, None].expand(-1, -1, xc.shape[2]))
        dcc = 1.0 - torch.bmm(xc, xc.transpose(1, 2)).float()    # [m,C,C]
        alive = dpc < 1e8
        # duplicates (same id twice) : keep the first
        same = (c[:, :, None] == c[:, None, :]) & torch.tril(torch.ones(C, C, dtype=torch.bool, device=x.device), -1)[None]
        alive &= ~same.any(dim=2)
        count = torch.zeros(e - s, dtype=torch.int64, device=x.device)
        sel = torch.zeros_like(alive)
        for i in range(C):
            take = alive[:, i] & (count < R)
            sel[:, i] = take
            count += take
            occl = (alpha * dcc[:, i, :] <= dpc) & take[:, None]
            occl[:, : i + 1] = False
            alive &= ~occl
        pos = torch.cumsum(sel, dim=1) - 1
        rows = torch.arange(e - s, device=x.device)[:, None].expand(-1, C)
        o = out[s:e]
        o[rows[sel], pos[sel]] = c[sel]
    return out


@torch.no_grad()
def add_reverse_edges(nbrs: torch.Tensor, cap: int):
    """nbrs [n,R] (-1 pad) -> candidate lists [n,cap]: own edges + every reverse edge (deduplicated,
    truncated to `cap` arbitrary-but-deterministic entries)."""
    n, R = nbrs.shape
    src = torch.arange(n, device=nbrs.device)[:, None].expand(-1, R)
    m = nbrs >= 0
    a, b = src[m], nbrs[m]
    node = torch.cat([a, b])
    other = torch.cat([b, a])
    key = torch.unique(node * n + other)          # sorted, deduplicated
    node, other = key // n, key % n
    start = torch.searchsorted(node, torch.arange(n, device=nbrs.device))
    pos = torch.arange(node.numel(), device=nbrs.device) - start[node]
    keep = pos < cap
    out = torch.full((n, cap), -1, dtype=torch.int64, device=nbrs.device)
    out[node[keep], pos[keep]] = other[keep]
    return out


@torch.no_grad()
def gen_labels(n: int, seed: int, max_label: int = 16, device="cuda"):
    """1-2 labels uniform in 1..max_label per node (mirrors build.rs:1988-1991). -> label_off u32[n+1], labels i16"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    a = torch.randint(1, max_label + 1, (n,), generator=g, device=device)
    b = torch.randint(1, max_label + 1, (n,), generator=g, device=device)
    two = torch.randint(0, 2, (n,), generator=g, device=device).bool() & (a != b)
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    cnt = 1 + two.long()
    off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(cnt, 0)
    labels = torch.empty(int(off[-1]), dtype=torch.int16, device=device)
    first = torch.where(two, lo, a)
    labels[off[:-1]] = first.to(torch.int16)
    labels[off[:-1][two] + 1] = hi[two].to(torch.int16)
    return off.to(torch.int32).cpu().numpy().astype(np.uint32), labels.cpu().numpy()


@torch.no_grad()
def quantize_nodes(x: torch.Tensor, distance_type: int, bits: int, mean: np.ndarray, m2: np.ndarray, count: int,
                   chunk: int = 65536) -> np.ndarray:
    """SBQ codes of every row through the product's quantizer kernel (dann_prepare_queries)."""
    from pgvectorscale_b200 import diskann
    n, dim = x.shape
    words = code_words(dim, bits)
    empty = Snapshot(n=0, dim=dim, dim_index=dim, bits=bits, words=words, R=8, distance_type=distance_type,
                     has_labels=False, count=count, mean=mean, m2=m2,
                     codes=np.zeros((0, words), np.uint64), nbrs=np.zeros((0, 8), np.uint32),
                     heap_tid=np.zeros(0, np.uint64), vectors=np.zeros((0, dim), np.float32))
    codes = np.empty((n, words), np.uint64)
    with diskann.DiskAnnIndex(empty, device=x.device.index or 0) as q:
        cw = q.code_stride
        buf = torch.empty((chunk, cw), dtype=torch.int64, device=x.device)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            q.prepare_queries(x[s:e].contiguous(), None, buf[: e - s])
            codes[s:e] = buf[: e - s, :words].cpu().numpy().view(np.uint64)
    return codes


@torch.no_grad()
def build_index(x: torch.Tensor, distance_type: int = COSINE, bits: int | None = None, R: int = 50,
                knn: int = 64, alpha: float = 1.2, labels_seed: int | None = None, log=None) -> Snapshot:
    """x: [n,dim] unit-norm f32 on the GPU -> host Snapshot (codes, graph, tids, vectors)."""
    t0 = time.time()
    n, dim = x.shape
    bits = default_bits(dim) if bits is None else bits
    say = log or (lambda *a: None)
    mean = x.mean(dim=0)
    m2 = ((x - mean) ** 2).sum(dim=0)
    mean_h, m2_h = mean.cpu().numpy().astype(np.float32), m2.cpu().numpy().astype(np.float32)
    codes = quantize_nodes(x, distance_type, bits, mean_h, m2_h, n)
    say(f"  sbq codes {time.time() - t0:.1f}s")
    cand, _ = knn_candidates(x, knn)
    say(f"  knn {time.time() - t0:.1f}s")
    g1 = alpha_prune(x, cand, R, alpha)
    del cand
    say(f"  prune1 {time.time() - t0:.1f}s")
    cand2 = add_reverse_edges(g1, 2 * knn)
    del g1
    g2 = alpha_prune(x, cand2, R, alpha, chunk=1024)
    del cand2
    say(f"  reverse+prune2 {time.time() - t0:.1f}s")
    nb = g2.cpu().numpy()
    nbrs = np.where(nb >= 0, nb, INVALID_NODE).astype(np.uint32)
    # entry point: the node closest to the centroid (the reference uses the first inserted node)
    start = int(torch.argmax(x @ (mean / mean.norm().clamp(min=1e-12))).item()) if n else INVALID_NODE
    label_off = labels = sl = sn = None
    if labels_seed is not None:
        label_off, labels = gen_labels(n, labels_seed, device=x.device)
        # per-label start node = first node (in heap order) carrying the label (graph/mod.rs:490-533)
        node_of = np.repeat(np.arange(n, dtype=np.uint32), np.diff(label_off).astype(np.int64))
        sl = np.unique(labels)
        first = {}
        for lab, nd in zip(labels[::-1].tolist(), node_of[::-1].tolist()):
            first[lab] = nd
        sn = np.array([first[int(l)] for l in sl], np.uint32)
    snap = Snapshot(n=n, dim=dim, dim_index=dim, bits=bits, words=code_words(dim, bits), R=R,
                    distance_type=distance_type, has_labels=labels_seed is not None, count=n, mean=mean_h,
                    m2=m2_h, codes=codes, nbrs=nbrs, heap_tid=make_heap_tids(n), vectors=x.cpu().numpy(),
                    start_default=start, start_labels=sl, start_label_nodes=sn, label_off=label_off,
                    labels=labels)
    say(f"  snapshot {time.time() - t0:.1f}s")
    return snap


@torch.no_grad()
def ground_truth(x: torch.Tensor, q: torch.Tensor, k: int, mask: torch.Tensor | None = None, chunk: int = 256):
    """Exact top-k node ids by cosine in f32 (recall ground truth only). mask [B,n] bool optional."""
    out = torch.empty((q.shape[0], k), dtype=torch.int64, device=x.device)
    torch.backends.cuda.matmul.allow_tf32 = False
    for s in range(0, q.shape[0], chunk):
        e = min(q.shape[0], s + chunk)
        sc = q[s:e] @ x.T
        if mask is not None:
            sc = torch.where(mask[s:e], sc, torch.full_like(sc, -4.0))
        out[s:e] = torch.topk(sc, k, dim=1).indices
    return out


@torch.no_grad()
def label_start_nodes(label_off: np.ndarray, labels: np.ndarray):
    """First node (in heap order) carrying each label: graph/mod.rs:490-533 update_start_nodes."""
    n = len(label_off) - 1
    node_of = np.repeat(np.arange(n, dtype=np.uint32), np.diff(label_off).astype(np.int64))
    sl, first_idx = np.unique(labels, return_index=True)
    return sl.astype(np.int16), node_of[first_idx].astype(np.uint32)


@torch.no_grad()
def build_index_vamana(x: torch.Tensor, distance_type: int = COSINE, bits: int | None = None, R: int = 50,
                       L_build: int = 100, alpha: float = 1.2, max_batch: int = 1 << 20, log=None, keep_index=False,
                       labels_seed: int | None = None):
    """Same contract as build_index(), but the graph comes from the product's GPU batch Vamana builder
    (dann_build_graph: beam search in build mode + alpha prune + back-links, all over SBQ codes, like the
    reference's build) instead of the exact-kNN fixture.  Scales linearly in n.
    Returns the host Snapshot (R = 64 slots per node, lists of <= R ids); with keep_index=True also the
    loaded DiskAnnIndex (vectors already supplied)."""
    from pgvectorscale_b200 import diskann
    t0 = time.time()
    n, dim = x.shape
    bits = default_bits(dim) if bits is None else bits
    say = log or (lambda *a: None)
    mean = x.mean(dim=0)
    m2 = ((x - mean) ** 2).sum(dim=0)
    mean_h, m2_h = mean.cpu().numpy().astype(np.float32), m2.cpu().numpy().astype(np.float32)
    codes = quantize_nodes(x, distance_type, bits, mean_h, m2_h, n)
    say(f"  sbq codes {time.time() - t0:.1f}s")
    slots = 64
    label_off = labels = sl = sn = None
    if labels_seed is not None:
        label_off, labels = gen_labels(n, labels_seed, device=x.device)
        sl, sn = label_start_nodes(label_off, labels)
    snap = Snapshot(n=n, dim=dim, dim_index=dim, bits=bits, words=code_words(dim, bits), R=slots,
                    distance_type=distance_type, has_labels=labels_seed is not None, count=n, mean=mean_h, m2=m2_h,
                    codes=codes, nbrs=np.full((n, slots), INVALID_NODE, np.uint32), heap_tid=make_heap_tids(n),
                    vectors=None, start_default=0 if n else INVALID_NODE, start_labels=sl, start_label_nodes=sn,
                    label_off=label_off, labels=labels)
    idx = diskann.DiskAnnIndex(snap, device=x.device.index or 0)
    st = idx.build_graph(R, L_build, alpha, max_batch)
    say(f"  gpu vamana build {time.time() - t0:.1f}s: {st}")
    snap.nbrs = idx.download_nbrs()
    snap.vectors = x.cpu().numpy()
    if keep_index:
        idx.set_vectors(snap.vectors)
        return snap, idx, st
    idx.close()
    return snap, None, st
