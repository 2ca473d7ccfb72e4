"""GPU batch Vamana builder: structural validity of the graph, recall of scans
over it, and scan parity (oracle vs CUDA on the SAME built snapshot)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gpu_built_graph_is_valid_searchable_and_parity_holds():
    import torch
    from oracle import oracle
    from pgvectorscale_b200 import diskann
    from tools import synth_index as si
    if diskann.device_count() < 1:
        pytest.fail("no CUDA device visible")
    dev = torch.device("cuda", 0)
    n, dim, R = 60_000, 256, 50
    x = si.gen_dataset(n, dim, 123, "lowrank", device=dev)
    snap, idx, st = si.build_index_vamana(x, R=R, L_build=100, keep_index=True)
    try:
        nb = snap.nbrs
        assert nb.shape == (n, 64)
        valid = nb != 0xFFFFFFFF
        deg = valid.sum(1)
        assert deg.max() <= R and deg[1:].min() >= 1 and st["batches"] > 10
        # lists are INVALID-terminated prefixes, ids in range, no self loops, no duplicates
        assert (valid[:, :-1] >= valid[:, 1:]).all()
        assert (nb[valid] < n).all()
        assert not (nb == np.arange(n, dtype=np.uint32)[:, None]).any()
        srt = np.sort(np.where(valid, nb, np.arange(n, dtype=np.uint32)[:, None] + np.uint32(2**31)), axis=1)
        assert not ((srt[:, 1:] == srt[:, :-1]) & (srt[:, 1:] < n)).any()
        assert deg.mean() > 0.6 * R
        # scans over the built graph
        q = si.gen_dataset(256, dim, 321, "lowrank", device=dev)
        truth = si.ground_truth(x, q, 10).cpu().numpy()
        qh = q.cpu().numpy()
        g = idx.search_batch(qh, k=10, search_list_size=100, rescore=100)
        tid = g["tid"]
        nodes = (tid >> np.uint64(16)).astype(np.int64) * 2 + (tid & np.uint64(0xFFFF)).astype(np.int64) - 1
        rec = np.mean([len(set(nodes[i].tolist()) & set(truth[i].tolist())) / 10 for i in range(len(nodes))])
        assert rec >= 0.9, rec
        otid, odist, _, ostats = oracle.scan_batch(snap, qh[:48], None, None, 100, 100, 10)
        assert np.array_equal(tid[:48], otid)
        assert np.array_equal(g["dist"][:48].view(np.uint32), odist.view(np.uint32))
        assert np.array_equal(g["stats"]["visits"][:48].astype(np.uint64), ostats["visits"])
    finally:
        idx.close()
