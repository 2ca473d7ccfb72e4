"""QPS @ recall at scales the exact-kNN fixture cannot reach (10M-50M x 768-d on one B200).

Data is generated chunk by chunk on the GPU (same "Cohere-shape" generator as bench.py), SBQ codes go to
HBM, the graph is built by the product's GPU batch Vamana builder (dann_build_graph),
then the f32 vectors are uploaded for the rerank.  Ground truth = exact f32 brute force accumulated
per chunk.  Reports the recall/QPS sweep, the first operating point with recall@10 >= target, and
oracle parity on a sample at that point.
   python tools/large_recall.py --n 50000000 --batch 4096 > profiles/r01_50m_recall.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import COSINE, INVALID_NODE, Snapshot, code_words, make_heap_tids
from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--bits", type=int, default=2)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--R", type=int, default=50)
ap.add_argument("--L-build", type=int, default=100)
ap.add_argument("--max-batch", type=int, default=1 << 20)
ap.add_argument("--target", type=float, default=0.99)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--cpu-sample", type=int, default=256)
ap.add_argument("--labels", action="store_true", help="16-way labels (1-2 per node), every query filters on one label")
a = ap.parse_args()
dev = torch.device("cuda", 0)
t0 = time.time()
n, dim, B, k = a.n, a.dim, a.batch, 10
words = code_words(dim, a.bits)
chunk = 1 << 20
q = si.gen_dataset(B, dim, 0x5EED0031, "lowrank", device=dev)
vec = np.empty((n, dim), np.float32)
mean = torch.zeros(dim, device=dev, dtype=torch.float64)
sq = torch.zeros(dim, device=dev, dtype=torch.float64)
label_off = labels = sl = sn = node_bits = None
qkeys = None
if a.labels:
    label_off, labels = si.gen_labels(n, 0x5EED0040, device=dev)
    sl, sn = si.label_start_nodes(label_off, labels)
    node_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(label_off).astype(np.int64))
    bits_h = np.zeros(n, np.int32)
    np.bitwise_or.at(bits_h, node_of, (1 << labels.astype(np.int32)))
    node_bits = torch.from_numpy(bits_h).to(dev)
    del node_of, bits_h
    qkeys = [[1 + (b % 16)] for b in range(B)]
    qbit = torch.tensor([1 << kk[0] for kk in qkeys], dtype=torch.int32, device=dev)
best_s = torch.full((B, k), -4.0, device=dev)
best_i = torch.full((B, k), -1, device=dev, dtype=torch.int64)
torch.backends.cuda.matmul.allow_tf32 = False
for s in range(0, n, chunk):
    e = min(n, s + chunk)
    x = si.gen_dataset(e - s, dim, 0x5EED0030 + s // chunk, "lowrank", device=dev)
    vec[s:e] = x.cpu().numpy()
    mean += x.double().sum(0)
    sq += (x.double() ** 2).sum(0)
    sc = q @ x.T
    if a.labels:
        sc = torch.where((node_bits[s:e][None, :] & qbit[:, None]) != 0, sc, torch.full_like(sc, -4.0))
    v, i = torch.topk(sc, min(k, e - s), dim=1)
    cs = torch.cat([best_s, v], 1)
    ci = torch.cat([best_i, i + s], 1)
    o = torch.topk(cs, k, dim=1).indices
    best_s, best_i = torch.gather(cs, 1, o), torch.gather(ci, 1, o)
    del x, sc
mean /= n
m2 = sq - n * mean ** 2
mean_h, m2_h = mean.float().cpu().numpy(), m2.float().cpu().numpy()
truth = best_i.cpu().numpy()
codes = np.empty((n, words), np.uint64)
for s in range(0, n, 4 * chunk):
    e = min(n, s + 4 * chunk)
    codes[s:e] = si.quantize_nodes(torch.from_numpy(vec[s:e]).to(dev), COSINE, a.bits, mean_h, m2_h, n)
qh = q.cpu().numpy()
del q, best_s, best_i
if a.labels:
    del node_bits
torch.cuda.empty_cache()
t_data = time.time() - t0
snap = Snapshot(n=n, dim=dim, dim_index=dim, bits=a.bits, words=words, R=64, distance_type=COSINE,
                has_labels=a.labels, count=n, mean=mean_h, m2=m2_h, codes=codes,
                nbrs=np.full((n, 64), INVALID_NODE, np.uint32), heap_tid=make_heap_tids(n), vectors=None,
                start_default=0, start_labels=sl, start_label_nodes=sn, label_off=label_off, labels=labels)
idx = diskann.DiskAnnIndex(snap)
t1 = time.time()
bst = idx.build_graph(a.R, a.L_build, 1.2, a.max_batch)
t_build = time.time() - t1
snap.nbrs = idx.download_nbrs()
t2 = time.time()
idx.set_vectors(vec)
snap.vectors = vec
t_vec = time.time() - t2
deg = (snap.nbrs[: 1 << 20] != INVALID_NODE).sum(1)
res = {"labels": bool(a.labels), "n": n, "dim": dim, "bits": a.bits, "batch": B, "R": a.R, "L_build": a.L_build, "data_s": round(t_data, 1),
       "build_s": round(t_build, 1), "build": {k2: (round(v2, 1) if isinstance(v2, float) else v2) for k2, v2 in bst.items()},
       "vectors_upload_s": round(t_vec, 1), "hbm_gb": round(idx.hbm_bytes / 1e9, 2),
       "degree_mean": float(deg.mean()), "degree_min": int(deg[1:].min()), "sweep": []}
print(f"[large_recall] data {t_data:.0f}s build {t_build:.0f}s {bst}", file=sys.stderr, flush=True)


def nodes_of(tid):
    nd = (tid >> np.uint64(16)).astype(np.int64) * 2 + (tid & np.uint64(0xFFFF)).astype(np.int64) - 1
    nd[tid == np.uint64(0xFFFFFFFFFFFFFFFF)] = -1
    return nd


SWEEP = [(50, 50), (100, 50), (100, 100), (100, 150), (100, 200), (150, 200), (200, 200), (200, 300), (300, 300),
         (400, 400), (600, 600), (800, 800), (1000, 1000), (1500, 1000)]
chosen = None
for (L, rescore) in SWEEP:
    idx.search_batch(qh, labels=qkeys, k=k, search_list_size=L, rescore=rescore)
    g = idx.search_batch(qh, labels=qkeys, k=k, search_list_size=L, rescore=rescore)
    t = idx.last_batch_timing()
    nd = nodes_of(g["tid"])
    rec = float(np.mean([len(set(nd[i].tolist()) & set(truth[i].tolist())) / k for i in range(B)]))
    row = {"L": L, "rescore": rescore, "recall": round(rec, 4), "device_ms": round(t["total_ms"], 3),
           "search_ms": round(t["search_ms"], 3), "qps": round(B / t["total_ms"] * 1e3),
           "visits": float(g["stats"]["visits"].mean()), "d_quantized": float(g["stats"]["d_quantized"].mean())}
    res["sweep"].append(row)
    print(f"[large_recall] {row}", file=sys.stderr, flush=True)
    if rec >= a.target and chosen is None:
        chosen = (L, rescore, rec, row["qps"])
        break
if chosen is None:
    chosen = (SWEEP[-1][0], SWEEP[-1][1], res["sweep"][-1]["recall"], res["sweep"][-1]["qps"])
L, rescore, rec, qps = chosen
res["operating_point"] = {"L": L, "rescore": rescore, "recall_at_10": round(rec, 4), "qps_device": qps}
from oracle import oracle
oracle.build_lib()
g = idx.search_batch(qh[:a.check], labels=qkeys[:a.check] if qkeys else None, k=k, search_list_size=L, rescore=rescore)
olab = ooff = None
if a.labels:
    olab = np.array([kk[0] for kk in qkeys[:a.check]], np.int16)
    ooff = np.arange(a.check + 1, dtype=np.int32)
otid, odist, _, ostats = oracle.scan_batch(snap, qh[:a.check], olab, ooff, L, rescore, k)
res["parity"] = {"queries": a.check, "tids_identical": bool(np.array_equal(g["tid"], otid)),
                 "dist_bits_identical": bool(np.array_equal(g["dist"].view(np.uint32), odist.view(np.uint32))),
                 "counters_identical": bool(np.array_equal(g["stats"]["visits"].astype(np.uint64), ostats["visits"]))}
ns = min(a.cpu_sample, B)
tc = time.perf_counter()
oracle.scan_batch(snap, qh[:ns], np.array([kk[0] for kk in qkeys[:ns]], np.int16) if qkeys else None,
                  np.arange(ns + 1, dtype=np.int32) if qkeys else None, L, rescore, k, threads=0)
res["cpu_baseline"] = {"qps": round(ns / (time.perf_counter() - tc), 1), "cores": os.cpu_count(), "sample": ns}
print(json.dumps(res))
