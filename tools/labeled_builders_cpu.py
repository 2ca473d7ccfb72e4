"""CPU-only: the reference's SERIAL labeled build (oracle) against the product's BATCH builder (dann_build_graph, run under
the emulated ABI - tests/simt) on the same rows, codes and labels: label-filtered recall of oracle scans over each graph.
Test infrastructure only (it loads the emulator's library by explicit path).

    python tools/labeled_builders_cpu.py [--n 6000 --dim 128] > profiles/r02_labeled_builders_emulated.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
import numpy as np

import build_emu
from oracle import fixtures, oracle
from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import COSINE, INVALID_NODE
from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=6000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--queries", type=int, default=200)
ap.add_argument("--max-batch", type=int, default=1 << 20)
a = ap.parse_args()
os.environ.setdefault("SIMT_SM_COUNT", "8")
diskann._LIB = diskann.load_library(build_emu.build_abi())
threads = max(1, len(os.sched_getaffinity(0)))
x = si.gen_dataset(a.n, a.dim, 1234, "lowrank", device="cpu").numpy()
q = si.gen_dataset(a.queries, a.dim, 99, "lowrank", device="cpu").numpy()
off, lab = fixtures.gen_labels(a.n, 77)
t0 = time.time()
serial = fixtures.make_index(x, COSINE, label_off=off, labels=lab)
t_serial = time.time() - t0
import copy
batch = copy.copy(serial)
batch.R = 64
batch.nbrs = np.full((a.n, 64), INVALID_NODE, np.uint32)
t0 = time.time()
with diskann.DiskAnnIndex(batch) as idx:
    st = idx.build_graph(50, 100, 1.2, a.max_batch)
    batch.nbrs = idx.download_nbrs()
t_batch = time.time() - t0
rng = np.random.default_rng(5)
keys = rng.integers(1, 17, size=a.queries).astype(np.int16)
qoff = np.arange(a.queries + 1, dtype=np.int32)
has = [set(lab[off[i]:off[i + 1]].tolist()) for i in range(a.n)]
sims = q @ x.T
truth = []
for b in range(a.queries):
    ok = np.fromiter((int(keys[b]) in has[i] for i in range(a.n)), dtype=bool, count=a.n)
    truth.append(set(np.argsort(-np.where(ok, sims[b], -2.0))[:10].tolist()))
lookup = {int(t): i for i, t in enumerate(serial.heap_tid)}
out = {"what": "label-filtered recall@10 of oracle scans over the serial (reference) graph and over the batch builder's graph "
               "(emulated ABI), same rows / codes / labels / queries", "n": a.n, "dim": a.dim, "queries": a.queries,
       "serial_build_seconds": round(t_serial, 1), "batch_build_seconds_emulated": round(t_batch, 1), "batch_build_stats": st, "graphs": {}}
for name, s in (("serial", serial), ("batch", batch)):
    rows = []
    for L, rescore in [(20, 20), (50, 50), (100, 50), (200, 200), (400, 400)]:
        tid, _, cnt, stt = oracle.scan_batch(s, q, keys, qoff, L, rescore, 10, threads=threads)
        hits = sum(len({lookup.get(int(t), -1) for t in tid[b][:cnt[b]]} & truth[b]) for b in range(a.queries))
        rows.append({"L": L, "rescore": rescore, "recall_at_10": round(hits / (a.queries * 10), 4), "visits": float(stt["visits"].mean())})
        print(name, rows[-1], file=sys.stderr, flush=True)
    nb = s.nbrs
    deg = (nb != INVALID_NODE).sum(1)
    same = [np.mean([len(has[i] & has[int(j)]) > 0 for j in nb[i][nb[i] != INVALID_NODE]]) for i in range(0, a.n, max(1, a.n // 1500)) if deg[i]]
    out["graphs"][name] = {"mean_degree": float(deg.mean()), "neighbours_sharing_a_label": round(float(np.mean(same)), 3), "sweep": rows}
print(json.dumps(out, indent=1))
