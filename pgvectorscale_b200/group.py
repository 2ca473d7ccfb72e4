"""Query-batch data parallelism over replicated indexes.

Queries are independent units: every rank holds a full replica of the index in its own HBM,
takes a contiguous slice of the batch, searches it locally, and the only exchange is one
gather of `B/N x k` (tid u64, dist f32) rows at the end.  One process per GPU;
`torch.distributed` is the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `total` queries for `rank` (first ranks take the remainder)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class QueryShardGroup:
    """Splits a global query batch across the ranks of a process group and gathers the top-k rows.

    search_fn(queries_slice) -> (tid [b,k] int64, dist [b,k] float32) tensors on `device`
    (for the product it wraps DiskAnnIndex.search_batch_device on this rank's replica)."""

    def __init__(self, search_fn: Callable, k: int, device, group: Optional[dist.ProcessGroup] = None):
        self.search_fn = search_fn
        self.k = k
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def gather_packed(self, tid: torch.Tensor, dst_dist: torch.Tensor, per_rank: int):
        """ONE collective per step: (tid, dist bits) packed into a single int64 block per rank and all_gathered on the
        CURRENT stream's order (torch makes the collective wait for the current stream and the current stream wait
        for the collective), so a search launched on that stream afterwards cannot overwrite `tid` under it.
        Returns (tids [world*per_rank, k] int64, dists [world*per_rank, k] float32)."""
        if self.world == 1:
            return tid, dst_dist
        k = self.k
        pack = torch.empty((per_rank, 2 * k), dtype=torch.int64, device=self.device)
        pack[:, :k] = tid
        pack[:, k:] = dst_dist.contiguous().view(torch.int32).to(torch.int64)
        out = torch.empty((self.world * per_rank, 2 * k), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, pack, group=self.group)
        return out[:, :k].contiguous(), out[:, k:].to(torch.int32).view(torch.float32)

    def search(self, queries: torch.Tensor):
        """queries: the GLOBAL batch [B, dim] (same tensor on every rank). Returns the global
        (tid [B,k], dist [B,k]) on every rank, rows in query order."""
        B = queries.shape[0]
        lo, hi = shard_bounds(B, self.world, self.rank)
        per_rank = (B + self.world - 1) // self.world
        tid = torch.full((per_rank, self.k), -1, dtype=torch.int64, device=self.device)
        dst = torch.full((per_rank, self.k), float("nan"), dtype=torch.float32, device=self.device)
        if hi > lo:
            t, d = self.search_fn(queries[lo:hi])
            tid[: hi - lo] = t
            dst[: hi - lo] = d
        tids, dists = self.gather_packed(tid, dst, per_rank)
        if self.world == 1:
            return tids[:B], dists[:B]
        rows = []
        for r in range(self.world):
            a, b = shard_bounds(B, self.world, r)
            rows.append(torch.arange(r * per_rank, r * per_rank + (b - a), device=self.device))
        sel = torch.cat(rows)
        return tids[sel], dists[sel]
