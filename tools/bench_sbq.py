"""Stand-alone SBQ-distance kernel timing (the kernel BASELINE.json's metric names).
   python tools/bench_sbq.py [--n 1000000] [--bits 2]   (knobs: DANN_SBQ_UNR / _THREADS / _BLOCKS_PER_SM)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import Snapshot, code_words

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--bits", type=int, default=2)
ap.add_argument("--pairs", type=int, default=64 * 1024 * 1024)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
words = code_words(a.dim, a.bits)
rng = np.random.default_rng(0)
codes = rng.integers(0, 2**63, size=(a.n, words), dtype=np.uint64)
s = Snapshot(n=a.n, dim=a.dim, dim_index=a.dim, bits=a.bits, words=words, R=8, distance_type=1, has_labels=False,
             count=a.n, mean=np.zeros(a.dim, np.float32), m2=np.ones(a.dim, np.float32), codes=codes,
             nbrs=np.full((a.n, 8), 0xFFFFFFFF, np.uint32), heap_tid=np.arange(a.n, dtype=np.uint64) * 65536 + 1,
             vectors=np.zeros((a.n, a.dim), np.float32), start_default=0)
idx = diskann.DiskAnnIndex(s)
g = torch.Generator(device=dev)
g.manual_seed(1)
pn = torch.randint(0, a.n, (a.pairs,), generator=g, device=dev, dtype=torch.int32)
pq = torch.randint(0, a.queries, (a.pairs,), generator=g, device=dev, dtype=torch.int32)
qc = torch.randint(0, 2**62, (a.queries, idx.code_stride), generator=g, device=dev, dtype=torch.int64)
out = torch.empty(a.pairs, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(dev)
for _ in range(3):
    idx.sbq_distance(qc, pq, pn, out, stream=st.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(st)
for _ in range(a.reps):
    idx.sbq_distance(qc, pq, pn, out, stream=st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
by = a.pairs * (idx.code_stride * 8 + 12)
# spot check
h = out[:1000].cpu().numpy()
ref = np.unpackbits((codes[pn[:1000].cpu().numpy()] ^ qc[pq[:1000].cpu().numpy().astype(np.int64), :words].cpu().numpy().view(np.uint64)).view(np.uint8), axis=1).sum(1)
peak = 6575.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps(dict(bits=a.bits, ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / peak, 4),
                      ok=bool((h == ref).all()), unr=os.environ.get("DANN_SBQ_UNR"), threads=os.environ.get("DANN_SBQ_THREADS"),
                      bps=os.environ.get("DANN_SBQ_BLOCKS_PER_SM"))))
