"""Parity + throughput at a scale the kNN fixture builder cannot reach (default 8M x 768-d).

The graph is a random duplicate-free R-regular digraph and the SBQ codes come from real
(low-rank synthetic) vectors, so recall is meaningless here; the point is the scan path itself at
TLB-scale tables: identical TIDs / distances / counters vs the oracle on the same snapshot, for both
inserted-set flavours (bitmap, CAS hash set), and the step time.
   python tools/scale_check.py --n 8000000 > profiles/r01_scale_check_8m.json"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import COSINE, Snapshot, code_words, make_heap_tids
from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--bits", type=int, default=2)
ap.add_argument("--R", type=int, default=50)
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--rescore", type=int, default=50)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--check", type=int, default=48)
ap.add_argument("--labels", action="store_true", help="16-way labels; every query filters on one label")
a = ap.parse_args()
dev = torch.device("cuda", 0)
t0 = time.time()
n, dim, R = a.n, a.dim, a.R
words = code_words(dim, a.bits)
vec = np.empty((n, dim), np.float32)
chunk = 1 << 20
mean = torch.zeros(dim, device=dev, dtype=torch.float64)
m2 = torch.zeros(dim, device=dev, dtype=torch.float64)
for s in range(0, n, chunk):
    e = min(n, s + chunk)
    x = si.gen_dataset(e - s, dim, 0x5EED0100 + s // chunk, "lowrank", device=dev)
    vec[s:e] = x.cpu().numpy()
    mean += x.double().sum(0)
    m2 += (x.double() ** 2).sum(0)
mean /= n
m2 = m2 - n * mean ** 2
mean_h, m2_h = mean.float().cpu().numpy(), m2.float().cpu().numpy()
codes = np.empty((n, words), np.uint64)
for s in range(0, n, 4 * chunk):
    e = min(n, s + 4 * chunk)
    codes[s:e] = si.quantize_nodes(torch.from_numpy(vec[s:e]).to(dev), COSINE, a.bits, mean_h, m2_h, n)
# random duplicate-free neighbour lists: nbr[i][j] = (i + 1 + (h_i + j*S) mod (n-1)) mod n, gcd(S, n-1) = 1
S = 104729
while math.gcd(S, n - 1) != 1:
    S += 2
nbrs = np.empty((n, R), np.uint32)
j = torch.arange(R, device=dev, dtype=torch.int64)
for s0 in range(0, n, 4 * chunk):
    e0 = min(n, s0 + 4 * chunk)
    i = torch.arange(s0, e0, device=dev, dtype=torch.int64)
    h = (i * 2654435761) % (n - 1)
    nbrs[s0:e0] = ((i[:, None] + 1 + (h[:, None] + j[None, :] * S) % (n - 1)) % n).to(torch.int32).cpu().numpy().view(np.uint32)
    del i, h
label_off = labels = sl = sn = None
if a.labels:
    label_off, labels = si.gen_labels(n, 0x5EED0040, device=dev)
    node_of = np.repeat(np.arange(n, dtype=np.uint32), np.diff(label_off).astype(np.int64))
    sl, first_idx = np.unique(labels, return_index=True)      # first node (heap order) carrying each label
    sn = node_of[first_idx].astype(np.uint32)
torch.cuda.empty_cache()
snap = Snapshot(n=n, dim=dim, dim_index=dim, bits=a.bits, words=words, R=R, distance_type=COSINE,
                has_labels=a.labels, count=n, mean=mean_h, m2=m2_h, codes=codes, nbrs=nbrs, heap_tid=make_heap_tids(n),
                vectors=vec, start_default=0, start_labels=sl, start_label_nodes=sn, label_off=label_off, labels=labels)
t_build = time.time() - t0
q = si.gen_dataset(4 * a.batch, dim, 0x5EED0200, "lowrank", device=dev).cpu().numpy()
torch.cuda.empty_cache()
qlab = [[1 + (b % 16)] for b in range(4 * a.batch)] if a.labels else None
idx = diskann.DiskAnnIndex(snap)
t_load = time.time() - t0 - t_build
res = {"labels": bool(a.labels), "n": n, "dim": dim, "bits": a.bits, "R": R, "L": a.L, "rescore": a.rescore, "batch": a.batch,
       "hbm_gb": round(idx.hbm_bytes / 1e9, 2), "fixture_s": round(t_build, 1), "load_s": round(t_load, 1)}
from oracle import oracle
olab = ooff = None
if a.labels:
    olab = np.array([l[0] for l in qlab[:a.check]], np.int16)
    ooff = np.arange(a.check + 1, dtype=np.int32)
otid, odist, _, ostats = oracle.scan_batch(snap, q[:a.check], olab, ooff, a.L, a.rescore, 10)
for name, env in (("bitmap", "1"), ("hash", "0")):
    os.environ["DANN_SEARCH_BITMAP"] = env
    ms = []
    for s in range(4):
        out = idx.search_batch(q[s * a.batch:(s + 1) * a.batch], labels=qlab[s * a.batch:(s + 1) * a.batch] if qlab else None,
                               k=10, search_list_size=a.L, rescore=a.rescore)
        ms.append(round(idx.last_batch_timing()["search_ms"], 3))
    g = idx.search_batch(q[:a.check], labels=qlab[:a.check] if qlab else None, k=10, search_list_size=a.L, rescore=a.rescore)
    res[name] = {"search_ms": ms, "retries": idx.last_batch_timing()["retries"],
                 "tids_identical": bool(np.array_equal(g["tid"], otid)),
                 "dist_bits_identical": bool(np.array_equal(g["dist"].view(np.uint32), odist.view(np.uint32))),
                 "counters_identical": bool(np.array_equal(g["stats"]["d_quantized"].astype(np.uint64), ostats["d_quantized"])
                                            and np.array_equal(g["stats"]["visits"].astype(np.uint64), ostats["visits"])),
                 "visits": float(out["stats"]["visits"].mean()), "d_quantized": float(out["stats"]["d_quantized"].mean())}
print(json.dumps(res))
