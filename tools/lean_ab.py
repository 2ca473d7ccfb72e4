"""A/B timing of the lean warp-per-query search kernel (dann_search3.cuh) against the round-1 two-warp kernel, and of
its push / pop engines and slot counts (DANN_SEARCH_KERNEL / DANN_HV_FLAGS / DANN_SEARCH_WARPS are read per call).

   python tools/make_snapshot.py --out /tmp/snap ; python tools/lean_ab.py --snap /tmp/snap --L 150 --rescore 250

One JSON line per variant: median device-timed search_ms over --steps batches, and parity of the first 32 queries
against the oracle (TIDs and rerank distance bits).  Numpy only (usable under ncu)."""
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
ap.add_argument("--batches", default="1024,4096")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
s = Snapshot.load(a.snap + ".npz")
q = np.load(a.snap + "_q.npy")
idx = diskann.DiskAnnIndex(s)
want = None
if not a.no_check:
    from oracle import oracle
    want = oracle.scan_batch(s, q[:32], None, None, a.L, a.rescore, a.k)

# (name, DANN_SEARCH_KERNEL, DANN_HV_FLAGS, DANN_SEARCH_WARPS, extra env)
VARIANTS = [("pairs (round 1)", 2, None, None, {}),
            ("lean, 4-level pop", 3, 2, None, {}),
            ("lean, lane-0 pop", 3, 0, None, {}),
            ("lean, staged pushes always", 3, 3, None, {}),
            ("lean (2) 20 warps/SM", 3, 2, 20, {}),
            ("lean (2) 24 warps/SM", 3, 2, 24, {}),
            ("lean (2) 32 warps/SM", 3, 2, 32, {}),
            ("lean (2) hash set", 3, 2, None, {"DANN_SEARCH_BITMAP": "0"})]
KEYS = ("DANN_SEARCH_KERNEL", "DANN_HV_FLAGS", "DANN_SEARCH_WARPS", "DANN_SEARCH_BITMAP", "DANN_SEARCH_ENTRY")
for B in [int(x) for x in a.batches.split(",")]:
    for name, kern, flags, warps, extra in VARIANTS:
        for k_ in KEYS:
            os.environ.pop(k_, None)
        os.environ["DANN_SEARCH_KERNEL"] = str(kern)
        if flags is not None:
            os.environ["DANN_HV_FLAGS"] = str(flags)
        if warps is not None:
            os.environ["DANN_SEARCH_WARPS"] = str(warps)
        os.environ.update(extra)
        if warps is not None and B < 148 * warps // 2:
            continue    # forcing more slots per SM than the batch fills only idles SMs
        ms, tot = [], []
        for i in range(a.steps + 2):
            qb = q[(i * B) % (len(q) - B + 1):][:B]
            idx.search_batch(qb, k=a.k, search_list_size=a.L, rescore=a.rescore)
            if i >= 2:
                t = idx.last_batch_timing()
                ms.append(t["search_ms"])
                tot.append(t.get("total_ms", 0.0))
        rec = dict(variant=name, batch=B, search_ms_median=round(float(np.median(ms)), 4), search_ms_min=round(float(np.min(ms)), 4),
                   qps_search_only=int(B / np.median(ms) * 1e3), total_ms_median=round(float(np.median(tot)), 4))
        if want is not None:
            g = idx.search_batch(q[:32], k=a.k, search_list_size=a.L, rescore=a.rescore)
            rec["parity"] = bool(np.array_equal(g["tid"], want[0]) and
                                 np.array_equal(g["dist"].view(np.uint32), want[1].view(np.uint32)))
        print(json.dumps(rec), flush=True)
