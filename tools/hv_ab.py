"""A/B timing of the HV=1 flavour of the two-warp search kernel against the measured default, one alternative at a
time (DANN_HEAP_V2 / DANN_HV_FLAGS are read per call, so one process and one loaded index serve every variant).

   python tools/make_snapshot.py ... ; python tools/hv_ab.py --snap /tmp/snap --L 150 --rescore 250

Prints one JSON line per variant: median device-timed search_ms over --steps batches, and parity of the first
32 queries against the oracle (TIDs and rerank distance bits).  Numpy only (usable under ncu)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pgvectorscale_b200 import diskann
from pgvectorscale_b200.snapshot import Snapshot

ap = argparse.ArgumentParser()
ap.add_argument("--snap", default="/tmp/snap")
ap.add_argument("--L", type=int, default=150)
ap.add_argument("--rescore", type=int, default=250)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
s = Snapshot.load(a.snap + ".npz")
q = np.load(a.snap + "_q.npy")
idx = diskann.DiskAnnIndex(s)
B = a.batch
want = None
if not a.no_check:
    from oracle import oracle
    want = oracle.scan_batch(s, q[:32], None, None, a.L, a.rescore, a.k)

VARIANTS = [("default", None, None), ("hv1 all", 1, 4095), ("hv1 none (same code paths as default)", 1, 0),
            ("push only", 1, 1), ("pop only", 1, 2), ("push+pop", 1, 3), ("distances only", 1, 4),
            ("code prefetch only", 1, 8), ("nbr prefetch only", 1, 16), ("visited search only", 1, 32),
            ("no intra-list dedupe only", 1, 64),
            ("root node id + prefetch by the heap warp only", 1, 128),
            ("heap v2 + root node (1+2+128)", 1, 131),
            ("TID prefetch only", 1, 256),
            ("fused expansion only", 1, 512),
            ("all but fused expansion", 1, 4095 - 512),
            ("uniform root prediction only", 1, 2048),
            ("heap warp all (1+2+128+2048)", 1, 2179),
            ("REDUX reductions only (with page-sized rounds: 4+1024)", 1, 1028),
            ("controller all (4+8+16+32+64+256+512)", 1, 892)]
# node-carrying entries are on for every hv1 variant above; one more line with them off
VARIANTS.append(("hv1 all, sequence-number entries (DANN_HV_NODE_ENTRIES=0)", 1, 4095))
for name, hv, flags in VARIANTS:
    os.environ["DANN_HV_NODE_ENTRIES"] = "0" if "DANN_HV_NODE_ENTRIES=0" in name else "1"
    for k_, v_ in (("DANN_HEAP_V2", hv), ("DANN_HV_FLAGS", flags)):
        if v_ is None:
            os.environ.pop(k_, None)
        else:
            os.environ[k_] = str(v_)
    ms = []
    for i in range(a.steps + 3):
        qb = q[(i * B) % (len(q) - B + 1):][:B]
        idx.search_batch(qb, k=a.k, search_list_size=a.L, rescore=a.rescore)
        if i >= 3:
            ms.append(idx.last_batch_timing()["search_ms"])
    rec = dict(variant=name, search_ms_median=round(float(np.median(ms)), 4), search_ms_min=round(float(np.min(ms)), 4))
    if want is not None:
        g = idx.search_batch(q[:32], k=a.k, search_list_size=a.L, rescore=a.rescore)
        rec["parity"] = bool(np.array_equal(g["tid"], want[0]) and
                             np.array_equal(g["dist"].view(np.uint32), want[1].view(np.uint32)))
    print(json.dumps(rec), flush=True)
