"""SURVEY §8e on the CPU: two gloo ranks, each with its own replica of the index behind the emulated ABI, split a query
batch with QueryShardGroup, search their slice with the REAL kernels (emulated) and gather the top-k - the global result
equals the oracle's, on every rank.  Emulated-ABI subprocess only (tests/test_emulated_abi.py); skipped on a GPU box,
where bench.py --gpus N exercises the same composition over NCCL."""
import os
import socket

import numpy as np
import pytest

from conftest import build_case, emulating

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not emulating(), reason="two emulated replicas; the GPU box uses bench.py --gpus N")]


def _worker(rank, world, port, out):
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "simt"))
    sys.path.insert(0, here)
    import build_emu
    from oracle import fixtures, oracle
    from pgvectorscale_b200 import diskann
    from pgvectorscale_b200.group import QueryShardGroup
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        diskann._LIB = None
        diskann._LIB = diskann.load_library(build_emu.build_abi())
        s = build_case(800, 64, 0, seed=33, kind="normal", R=24, L_build=48, deleted_every=7)   # same replica on every rank
        q = fixtures.gen_vectors(9, 64, 5, "normal")                                           # 9 queries over 2 ranks: 5 + 4
        with diskann.DiskAnnIndex(s) as idx:
            def search_fn(qs):
                g = idx.search_batch(qs.numpy(), k=10, search_list_size=40, rescore=20)
                return torch.from_numpy(g["tid"].view(np.int64)), torch.from_numpy(g["dist"])
            grp = QueryShardGroup(search_fn, k=10, device="cpu")
            tid, d = grp.search(torch.from_numpy(q))
        ok = True
        for b in range(len(q)):
            r = oracle.scan(s, q[b], None, 40, 20, 10)
            ok &= tid[b].numpy().view(np.uint64).tolist() == r["tid"].tolist()
            ok &= d[b].numpy().view(np.uint32).tolist() == r["dist"].view(np.uint32).tolist()
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_two_ranks_with_real_replicas_equal_the_oracle(lib_built):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == 1 and out[1] == 1


def test_graft_entry_smoke_under_emulation(lib_built, capsys):
    """__graft_entry__.smoke() - the driver's first call on the GPU box - against the emulated ABI."""
    import __graft_entry__ as g
    g.smoke()
    assert "smoke ok" in capsys.readouterr().out
