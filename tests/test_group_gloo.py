"""N>1 path on CPU: world_size-2 gloo run of the query-shard + gather logic (no GPU kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pgvectorscale_b200.group import QueryShardGroup, shard_bounds


def test_shard_bounds_cover_batch():
    for total in (0, 1, 7, 1024, 1025, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_search(q):
    # a deterministic stand-in for the per-rank replica search: rows are a function of the query only
    base = (q[:, 0] * 1000).round().to(torch.int64)
    tid = base[:, None] * 16 + torch.arange(4)[None, :]
    return tid, tid.to(torch.float32) * 0.5


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, 3) / 1000.0
        g = QueryShardGroup(_fake_search, k=4, device="cpu")
        tid, d = g.search(q)
        want_t, want_d = _fake_search(q)
        ok = torch.equal(tid, want_t) and torch.equal(d, want_d)
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 9, 1])
def test_two_rank_gather_matches_single_rank(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    assert out[0] == 1 and out[1] == 1
