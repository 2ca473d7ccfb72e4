"""Build the synthetic benchmark index once and save it (+ query batches) for profiling runs.
   python tools/make_snapshot.py --n 1000000 --out /tmp/snap"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tools import synth_index as si

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--data", default="lowrank")
ap.add_argument("--bits", type=int, default=0)
ap.add_argument("--queries", type=int, default=8192)
ap.add_argument("--labels", action="store_true")
ap.add_argument("--fixture", default="vamana", choices=["vamana", "knn"])
ap.add_argument("--out", default="/tmp/snap")
ap.add_argument("--raw", action="store_true", help="also write <out>.raw / <out>_q.f32 for the C harnesses")
a = ap.parse_args()
dev = torch.device("cuda", 0)
x = si.gen_dataset(a.n, a.dim, 0x5EED0010, a.data, device=dev)
if a.fixture == "vamana" and not a.labels:
    snap, _, _ = si.build_index_vamana(x, bits=a.bits or None, log=print)
else:
    snap = si.build_index(x, bits=a.bits or None, labels_seed=0x5EED0040 if a.labels else None, log=print)
q = si.gen_dataset(a.queries, a.dim, 0x5EED0011, a.data, device=dev)
truth = si.ground_truth(x, q[:1024], 10).cpu().numpy()
snap.save(a.out + ".npz")
np.save(a.out + "_q.npy", q.cpu().numpy())
np.save(a.out + "_truth.npy", truth)
if a.raw:
    snap.save_raw(a.out + ".raw")
    q.cpu().numpy().astype(np.float32).tofile(a.out + "_q.f32")
print("saved", a.out)
