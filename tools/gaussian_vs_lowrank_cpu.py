"""CPU-only evidence for the bench's choice of synthetic data: the reference's own algorithm
(the oracle's restatement: serial Vamana build over SBQ codes, streaming scan, exact rerank) on i.i.d. Gaussian rows versus
the "Cohere-shape" low-rank rows, same n / dim / graph parameters.  No GPU, no product code.

    python tools/gaussian_vs_lowrank_cpu.py [--n 20000] > profiles/r02_gaussian_vs_lowrank_oracle.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import fixtures, oracle
from pgvectorscale_b200.snapshot import COSINE
from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--queries", type=int, default=200)
a = ap.parse_args()
out = {"what": "recall@10 of the reference algorithm (CPU oracle, serial build, R=50, L_build=100, alpha=1.2, 2-bit SBQ, cosine) "
               "against exact brute force, per synthetic data kind", "n": a.n, "dim": a.dim, "queries": a.queries, "kinds": {}}
threads = max(1, len(os.sched_getaffinity(0)))
for kind in ("gaussian", "lowrank"):
    x = si.gen_dataset(a.n, a.dim, 1234, kind, device="cpu").numpy()
    q = si.gen_dataset(a.queries, a.dim, 99, kind, device="cpu").numpy()
    t0 = time.time()
    s = fixtures.make_index(x, COSINE)
    build_s = time.time() - t0
    truth = np.argsort(-(q @ x.T), axis=1)[:, :10]
    nn_cos = float(np.sort(q @ x.T, axis=1)[:, -1].mean())
    bulk = float((q @ x.T).std())
    rows = []
    for L, rescore in [(100, 50), (200, 200), (400, 400), (1000, 1000), (2000, 1000)]:
        tid, _, _, st = oracle.scan_batch(s, q, None, None, L, rescore, 10, threads=threads)
        node = ((tid >> np.uint64(16)) * np.uint64(0) + tid)       # heap tids of the fixture: node i -> make_heap_tids(n)[i]
        lookup = {int(t): i for i, t in enumerate(s.heap_tid)}
        hits = sum(len({lookup.get(int(t), -1) for t in tid[b]} & set(truth[b].tolist())) for b in range(a.queries))
        rows.append({"L": L, "rescore": rescore, "recall_at_10": round(hits / (a.queries * 10), 4),
                     "visits": float(st["visits"].mean()), "d_quantized": float(st["d_quantized"].mean())})
        print(kind, rows[-1], file=sys.stderr, flush=True)
    out["kinds"][kind] = {"nearest_neighbour_cosine_mean": round(nn_cos, 4), "cosine_std_over_the_dataset": round(bulk, 4),
                          "serial_build_seconds": round(build_s, 1), "sweep": rows}
print(json.dumps(out, indent=1))
