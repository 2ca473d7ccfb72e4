"""Timing / ncu driver at scales where the dataset must stay on the GPU (tools/fixture.py: rows generated straight into
the tensor the index borrows).  Builds the index with the product's GPU builder, then runs --steps batches at a fixed
operating point and prints per-step kernel timings, counters and the search plan.

   python tools/profile_big.py --n 8000000 --L 1000 --rescore 1000 --batch 4096 --steps 3
   ncu --set full -k regex:dann_search3_kernel -c 1 ... python tools/profile_big.py ...   (torch kernels are filtered by -k)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tools import fixture as fx

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--L", type=int, default=1000)
ap.add_argument("--rescore", type=int, default=1000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--check", type=int, default=0, help="compare this many queries with the oracle")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
t0 = time.time()
snap, idx, bst = fx.codes_and_graph(a.n, a.dim, "lowrank", 0, dev, log=lambda *m: print(*m, file=sys.stderr, flush=True),
                                    download_nbrs=a.check > 0)
X = torch.empty((a.n, a.dim), dtype=torch.float32, device=dev)
q = fx.gen_queries(a.batch, a.steps + 1, a.dim, "lowrank", dev)
topk = fx.RunningTopK(q[:a.batch], a.k)
fx.fill_rows(X, a.n, a.dim, "lowrank", dev, topk)
truth = topk.result()
idx.set_vectors_device(X.data_ptr())
torch.cuda.empty_cache()
print(f"[profile_big] fixture {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
B, k = a.batch, a.k
d_tid = torch.empty((B, k), dtype=torch.int64, device=dev)
d_dist = torch.empty((B, k), dtype=torch.float32, device=dev)
d_cnt = torch.empty(B, dtype=torch.int32, device=dev)
d_stats = torch.empty((B, 6), dtype=torch.int32, device=dev)
res = []
for s in range(a.steps + 1):
    idx.search_batch_device(q[s * B:(s + 1) * B], k, a.L, a.rescore, d_tid, d_dist, d_cnt, d_stats)
    torch.cuda.synchronize()
    t = idx.last_batch_timing()
    res.append({kk: round(v, 3) for kk, v in t.items()})
    if s == 0:
        nodes = fx.tid_to_node(d_tid.cpu().numpy().view(np.uint64))
        rec = float(np.mean([len(set(nodes[i].tolist()) & set(truth[i].tolist())) / k for i in range(B)]))
st = d_stats.cpu().numpy().astype(np.float64)
out = {"n": a.n, "L": a.L, "rescore": a.rescore, "batch": B, "recall_batch0": round(rec, 4), "steps": res,
       "visits": st[:, 0].mean(), "d_quantized": st[:, 1].mean(), "plan": idx.last_search_plan(),
       "alg_bytes_per_launch": float((st[:, 1].mean() * idx.code_stride * 8 + st[:, 0].mean() * 200) * B)}
out["search_gbs"] = round(out["alg_bytes_per_launch"] / (np.median([r["search_ms"] for r in res[1:]]) * 1e-3) / 1e9, 1)
if a.check:
    from oracle import oracle
    oracle.build_lib()
    rows = fx.SparseRows(a.n, a.dim)
    snap.vectors = rows.arr
    qh = q[:a.check].cpu().numpy()
    th = fx.host_cores()["effective"]
    rows.fill_from_device(fx.oracle_rerank_rows(oracle, snap, qh, a.L, a.rescore, k, th), X)
    idx.search_batch_device(q[:B], k, a.L, a.rescore, d_tid, d_dist, d_cnt, d_stats)
    torch.cuda.synchronize()
    otid, odist, _, _ = oracle.scan_batch(snap, qh, None, None, a.L, a.rescore, k, threads=th)
    out["parity"] = bool(np.array_equal(d_tid[:a.check].cpu().numpy().view(np.uint64), otid))
print(json.dumps(out))
