"""CPU-only look at label-filtered recall with the reference's own algorithm (oracle: serial labeled build, graph/mod.rs:637-660,
445-455; filtered scan, sbq/storage.rs:165-172) - the question config 4 (16-way label filter) raises: how much of the
filtered-recall deficit measured at 50M with the GPU-built graph (0.84 at L=1500) is the algorithm + uncorrelated labels,
and how much could be the batch builder.

    python tools/labeled_recall_cpu.py [--n 50000] > profiles/r02_labeled_recall_oracle.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import fixtures, oracle
from pgvectorscale_b200.snapshot import COSINE
from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--queries", type=int, default=200)
a = ap.parse_args()
threads = max(1, len(os.sched_getaffinity(0)))
x = si.gen_dataset(a.n, a.dim, 1234, "lowrank", device="cpu").numpy()
q = si.gen_dataset(a.queries, a.dim, 99, "lowrank", device="cpu").numpy()
off, lab = fixtures.gen_labels(a.n, 77)                       # 1-2 labels of 16 per node, uncorrelated with the geometry
t0 = time.time()
s = fixtures.make_index(x, COSINE, label_off=off, labels=lab)
build_s = time.time() - t0
rng = np.random.default_rng(5)
keys = rng.integers(1, 17, size=a.queries).astype(np.int16)
qoff = np.arange(a.queries + 1, dtype=np.int32)
has = [set(lab[off[i]:off[i + 1]].tolist()) for i in range(a.n)]
sims = q @ x.T
truth = []
for b in range(a.queries):
    ok = np.fromiter((int(keys[b]) in has[i] for i in range(a.n)), dtype=bool, count=a.n)
    sc = np.where(ok, sims[b], -2.0)
    truth.append(set(np.argsort(-sc)[:10].tolist()))
lookup = {int(t): i for i, t in enumerate(s.heap_tid)}
rows = []
for filt in (True, False):
    for L, rescore in [(100, 50), (200, 200), (400, 400), (1000, 1000), (2000, 1000)]:
        tid, _, cnt, st = oracle.scan_batch(s, q, keys if filt else None, qoff if filt else None, L, rescore, 10, threads=threads)
        if filt:
            hits = sum(len({lookup.get(int(t), -1) for t in tid[b][:cnt[b]]} & truth[b]) for b in range(a.queries))
            rec = hits / (a.queries * 10)
        else:
            un = [set(np.argsort(-sims[b])[:10].tolist()) for b in range(a.queries)]
            rec = sum(len({lookup.get(int(t), -1) for t in tid[b][:cnt[b]]} & un[b]) for b in range(a.queries)) / (a.queries * 10)
        rows.append({"filtered": filt, "L": L, "rescore": rescore, "recall_at_10": round(rec, 4), "visits": float(st["visits"].mean()),
                     "d_quantized": float(st["d_quantized"].mean())})
        print(rows[-1], file=sys.stderr, flush=True)
deg = (s.nbrs != 0xFFFFFFFF).sum(1)
same = []
for i in range(0, a.n, max(1, a.n // 2000)):
    nb = s.nbrs[i][s.nbrs[i] != 0xFFFFFFFF]
    same.append(np.mean([len(has[i] & has[int(j)]) > 0 for j in nb]) if len(nb) else 0.0)
print(json.dumps({"what": "label-filtered vs unfiltered recall@10 of the reference algorithm (CPU oracle, serial labeled build) on low-rank rows; "
                          "1-2 of 16 labels per node, one label per query", "n": a.n, "dim": a.dim, "queries": a.queries,
                  "serial_build_seconds": round(build_s, 1), "mean_degree": float(deg.mean()),
                  "fraction_of_a_nodes_neighbours_sharing_a_label_with_it": round(float(np.mean(same)), 3), "sweep": rows}, indent=1))
