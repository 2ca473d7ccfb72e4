"""Child process of `bench.py --impl reference`: builds the synthetic index fixture (SBQ codes + Vamana graph) with the
product's quantizer kernel and GPU batch builder - the reference's serial CPU build would take days at 50M - and
writes it to files, so that the process which times the reference's CPU algorithm never maps libdiskann_b200.so.

   python tools/make_fixture.py --n 50000000 --dim 768 --data lowrank --bits 0 --out /dev/shm/dann_fx
   -> <out>_codes.npy [n][words] u64, <out>_nbrs.npy [n][64] u32, <out>_meta.npz (mean, m2, bits, R, build stats)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tools import fixture as fx

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, required=True)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--data", default="lowrank")
ap.add_argument("--bits", type=int, default=0)
ap.add_argument("--device", type=int, default=0)
ap.add_argument("--out", required=True)
a = ap.parse_args()
dev = torch.device("cuda", a.device)
torch.cuda.set_device(dev)
t0 = time.time()
snap, idx, st = fx.codes_and_graph(a.n, a.dim, a.data, a.bits, dev, log=lambda *m: print(*m, file=sys.stderr, flush=True))
idx.close()
np.save(a.out + "_codes.npy", snap.codes)
np.save(a.out + "_nbrs.npy", snap.nbrs)
np.savez(a.out + "_meta.npz", mean=snap.mean, m2=snap.m2, bits=snap.bits, R=snap.R, words=snap.words,
         build=json.dumps({k: (float(v) if isinstance(v, float) else int(v)) for k, v in st.items()}))
print(f"[make_fixture] n={a.n} written to {a.out}_*.np[yz] in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
