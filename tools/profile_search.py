"""Numpy-only driver of the C ABI for ncu / timing runs (no torch: ncu sees only our kernels).
   python tools/profile_search.py --snap /tmp/snap --L 100 --rescore 50 --steps 5"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import Snapshot

ap = argparse.ArgumentParser()
ap.add_argument("--snap", default="/tmp/snap")
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--rescore", type=int, default=50)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--label", type=int, default=0, help="filter every query on this label (0 = no key)")
ap.add_argument("--check", action="store_true", help="compare the first 32 queries with the oracle")
a = ap.parse_args()
s = Snapshot.load(a.snap + ".npz")
q = np.load(a.snap + "_q.npy")
idx = diskann.DiskAnnIndex(s)
B = a.batch
labels = [[a.label]] * B if a.label else None
res = []
for i in range(a.steps):
    qb = q[(i * B) % (len(q) - B + 1):][:B]
    t0 = time.perf_counter()
    out = idx.search_batch(qb, labels=labels, k=a.k, search_list_size=a.L, rescore=a.rescore)
    wall = (time.perf_counter() - t0) * 1e3
    t = idx.last_batch_timing()
    res.append(dict(step=i, wall_ms=round(wall, 3), **{k: round(v, 4) for k, v in t.items()}))
st = out["stats"]
print(json.dumps(dict(L=a.L, rescore=a.rescore, B=B, steps=res, visits=float(st["visits"].mean()),
                      d_quantized=float(st["d_quantized"].mean()))))
if a.check:
    from oracle import oracle
    qb = q[:32]
    g = idx.search_batch(qb, labels=labels[:32] if labels else None, k=a.k, search_list_size=a.L, rescore=a.rescore)
    lab = off = None
    if labels:
        lab = np.full(32, a.label, np.int16)
        off = np.arange(33, dtype=np.int32)
    otid, odist, _, _ = oracle.scan_batch(s, qb, lab, off, a.L, a.rescore, a.k)
    print("parity:", bool(np.array_equal(g["tid"], otid)), bool(np.array_equal(g["dist"].view(np.uint32), odist.view(np.uint32))))
