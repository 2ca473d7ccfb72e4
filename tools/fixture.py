"""Benchmark fixtures at 1M-50M x 768-d, generated chunk by chunk on the GPU (torch) so that 50M rows (153.6 GB)
never exist twice: the rows are written straight into the tensor the index borrows (dann_index_set_vectors_device).

Shared by bench.py's two arms.  Only `codes_and_graph` touches the product library (SBQ quantizer kernel + GPU batch
Vamana builder); everything else is torch / numpy, so the `--impl reference` arm can call it from a child process
(tools/make_fixture.py) and never map libdiskann_b200.so itself.

Determinism: chunk c (CHUNK rows) of the dataset is gen_dataset(seed = DATA_SEED + c) - independent of n, so the 1M
index is a prefix of the 50M one and every process / rank regenerates identical rows.
"""
from __future__ import annotations

import mmap
import time

import numpy as np
import torch

from tools import synth_index as si

CHUNK = 1 << 18            # rows per generator call: 0.75 GiB per temporary, 50M rows leave ~3 GiB of HBM free
DATA_SEED = 0x5EED0030
QUERY_SEED0 = 0x5EED0031     # first batch of rank 0 (recall, operating point, parity)
QUERY_SEED = 0x5EED0033      # later batches


def chunks(n):
    for c, s in enumerate(range(0, n, CHUNK)):
        yield c, s, min(n, s + CHUNK)


def gen_chunk(c, rows, dim, data, device):
    return si.gen_dataset(rows, dim, DATA_SEED + c, data, device=device)


def gen_queries(B, nb, dim, data, device, rank=0):
    """[nb*B, dim]: batch 0 is generated on its own so that the other arm reproduces exactly the same queries."""
    parts = [si.gen_dataset(B, dim, QUERY_SEED0 + 7919 * rank, data, device=device)]
    if nb > 1:
        parts.append(si.gen_dataset((nb - 1) * B, dim, QUERY_SEED + 7919 * rank, data, device=device))
    return torch.cat(parts)


class RunningTopK:
    """Exact top-k by inner product (= cosine on unit rows) accumulated over dataset chunks, f32 (no TF32)."""

    def __init__(self, q: torch.Tensor, k: int):
        self.q, self.k = q, k
        self.best_s = torch.full((q.shape[0], k), -4.0, device=q.device)
        self.best_i = torch.full((q.shape[0], k), -1, device=q.device, dtype=torch.int64)
        torch.backends.cuda.matmul.allow_tf32 = False

    def add(self, x: torch.Tensor, start: int, sub: int = 1 << 16):
        for s in range(0, x.shape[0], sub):          # [B, sub] scores at a time: 4096 x 64K x 4 B = 1 GiB
            e = min(x.shape[0], s + sub)
            sc = self.q @ x[s:e].T
            v, i = torch.topk(sc, min(self.k, e - s), dim=1)
            cs = torch.cat([self.best_s, v], 1)
            ci = torch.cat([self.best_i, i + (start + s)], 1)
            o = torch.topk(cs, self.k, dim=1).indices
            self.best_s, self.best_i = torch.gather(cs, 1, o), torch.gather(ci, 1, o)
            del sc

    def result(self) -> np.ndarray:
        return self.best_i.cpu().numpy()


def dataset_stats(n, dim, data, device):
    """Per-dimension mean and m2 (sum of squared deviations) of the whole dataset in f64 -> f32 (SbqMeans)."""
    mean = torch.zeros(dim, device=device, dtype=torch.float64)
    sq = torch.zeros(dim, device=device, dtype=torch.float64)
    for c, s, e in chunks(n):
        x = gen_chunk(c, e - s, dim, data, device).double()
        mean += x.sum(0)
        sq += (x * x).sum(0)
        del x
    mean /= n
    m2 = sq - n * mean ** 2
    return mean.float().cpu().numpy(), m2.float().cpu().numpy()


def codes_and_graph(n, dim, data, bits, device, R=50, L_build=100, alpha=1.2, max_batch=1 << 20, log=None,
                    download_nbrs=True):
    """SBQ codes (product quantizer kernel) + graph (product GPU batch Vamana builder) of the synthetic dataset.
    -> (Snapshot with vectors=None [nbrs filled when download_nbrs], DiskAnnIndex still loaded, build stats)."""
    from pgvectorscale_b200 import diskann
    from pgvectorscale_b200.snapshot import COSINE, INVALID_NODE, Snapshot, code_words, default_bits, make_heap_tids
    say = log or (lambda *a: None)
    t0 = time.time()
    bits = bits or default_bits(dim)
    words = code_words(dim, bits)
    mean_h, m2_h = dataset_stats(n, dim, data, device)
    codes = np.empty((n, words), np.uint64)
    for c, s, e in chunks(n):
        x = gen_chunk(c, e - s, dim, data, device)
        codes[s:e] = si.quantize_nodes(x, COSINE, bits, mean_h, m2_h, n)
        del x
    say(f"  stats + sbq codes {time.time() - t0:.1f}s")
    slots = 64
    snap = Snapshot(n=n, dim=dim, dim_index=dim, bits=bits, words=words, R=slots, distance_type=COSINE,
                    has_labels=False, count=n, mean=mean_h, m2=m2_h, codes=codes,
                    nbrs=np.full((n, slots), INVALID_NODE, np.uint32), heap_tid=make_heap_tids(n), vectors=None,
                    start_default=0 if n else INVALID_NODE, start_labels=None, start_label_nodes=None,
                    label_off=None, labels=None)
    torch.cuda.empty_cache()
    idx = diskann.DiskAnnIndex(snap, device=device.index or 0)
    t1 = time.time()
    st = idx.build_graph(R, L_build, alpha, max_batch)
    say(f"  gpu vamana build {time.time() - t1:.1f}s: {st}")
    if download_nbrs:
        snap.nbrs = idx.download_nbrs()
    return snap, idx, st


def fill_rows(X: torch.Tensor, n, dim, data, device, topk: RunningTopK | None = None):
    """Writes the dataset into X (device, [n, dim]) chunk by chunk; feeds the exact-kNN accumulator on the way."""
    for c, s, e in chunks(n):
        x = gen_chunk(c, e - s, dim, data, device)
        X[s:e] = x
        if topk is not None:
            topk.add(x, s)
        del x


class SparseRows:
    """A [n, dim] f32 host array backed by an anonymous, lazily committed mapping: only the rows somebody writes take
    memory.  The CPU oracle reads heap vectors only for the rows it reranks, so a bounded query sample needs a few
    GB of the 153.6 GB table; a row that was not supplied reads as zeros, which changes the rerank distance and
    fails parity loudly rather than silently."""

    def __init__(self, n, dim):
        self.n, self.dim = n, dim
        nbytes = max(n * dim * 4, mmap.PAGESIZE)
        self.mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | getattr(mmap, "MAP_NORESERVE", 0))
        self.arr = np.frombuffer(self.mm, dtype=np.float32, count=n * dim).reshape(n, dim)
        self.have = np.zeros(n, dtype=bool)

    def missing(self, rows: np.ndarray) -> np.ndarray:
        rows = np.unique(rows.astype(np.int64))
        return rows[~self.have[rows]]

    def put(self, rows: np.ndarray, values: np.ndarray):
        self.arr[rows] = values
        self.have[rows] = True

    def fill_from_device(self, rows: np.ndarray, X: torch.Tensor, step: int = 1 << 18):
        rows = self.missing(rows)
        for s in range(0, len(rows), step):
            r = rows[s:s + step]
            self.put(r, X[torch.from_numpy(r).to(X.device)].cpu().numpy())

    def fill_from_generator(self, rows: np.ndarray, dim, data, device):
        """Regenerates the chunks that hold `rows` (for the arm that never keeps the dataset in HBM)."""
        rows = self.missing(rows)
        if not len(rows):
            return
        ch = rows // CHUNK
        for c in np.unique(ch):
            s = int(c) * CHUNK
            e = min(self.n, s + CHUNK)
            x = gen_chunk(int(c), e - s, dim, data, device)
            r = rows[ch == c]
            self.put(r, x[torch.from_numpy(r - s).to(device)].cpu().numpy())
            del x


def tid_to_node(tid: np.ndarray) -> np.ndarray:
    """Synthetic heap tids are a bijection of node ids (make_heap_tids): node = block * 2 + offset - 1; -1 = no row."""
    blk = (tid >> np.uint64(16)).astype(np.int64)
    off = (tid & np.uint64(0xFFFF)).astype(np.int64)
    nodes = blk * 2 + off - 1
    nodes[tid == np.uint64(0xFFFFFFFFFFFFFFFF)] = -1
    return nodes


def oracle_rerank_rows(oracle, snap, queries: np.ndarray, L: int, rescore: int, k: int, threads: int) -> np.ndarray:
    """The node ids whose heap vectors the reference algorithm fetches for these scans: k amgettuple calls consume the
    first rescore + k - 1 items of the approximate stream (scan.rs:255-305) - obtained from the oracle itself by
    scanning with rescore = 0 and k' = rescore + k - 1 (no heap vector is read without a rerank)."""
    if rescore == 0:
        return np.zeros(0, np.int64)
    c = rescore + k - 1
    tid, _, _, _ = oracle.scan_batch(snap, queries, None, None, L, 0, c, threads=threads)
    nodes = tid_to_node(tid).ravel()
    return np.unique(nodes[nodes >= 0])


def host_cores() -> dict:
    """CPUs this process may really use: the affinity mask AND the cgroup CPU quota (a 128-thread box with
    cpu.max = '1600000 100000' runs 16 cores' worth of threads)."""
    import math
    import os
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(p).read().split()
            if p.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read().split()[0])
            break
        except Exception:
            continue
    eff = aff if quota is None else max(1, min(aff, int(math.ceil(quota))))
    return {"logical": os.cpu_count() or 1, "affinity": aff, "cgroup_quota": quota, "effective": eff}
