"""Regenerates tests/golden/plain_golden.npz: frozen oracle answers for the plain storage layout (storage_layout =
plain).  Self-generated like scan_golden.npz (see make_golden.py for why): they keep later edits of
oracle.cpp from drifting and give the emulated / GPU kernels inputs with known answers.
    python tests/golden/make_plain_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: (n, dim, dim_index, distance, L, rescore, k)
    "plain_cos128": (500, 128, None, 0, 40, 20, 10),      # dim == dim_index: next(), no rerank
    "plain_l2_96x40": (500, 96, 40, 1, 30, 25, 12),       # truncated index slice: rerank from the heap vectors
    "plain_cos70x38": (400, 70, 38, 0, 25, 0, 8),         # odd widths, rescore = 0
}


def make_case(name):
    from conftest import build_case
    from oracle import fixtures
    n, dim, dim_index, dist, L, rescore, k = CASES[name]
    s = fixtures.to_plain(build_case(n, dim, dist, seed=303, kind="normal", R=24, L_build=48, deleted_every=11,
                                     dim_index=dim_index))
    q = fixtures.gen_vectors(6, dim, 404, "normal")
    return s, q, L, rescore, k


def run_case(name):
    from oracle import oracle
    s, q, L, rescore, k = make_case(name)
    tid = np.full((len(q), k), 0xFFFFFFFFFFFFFFFF, np.uint64)
    dist_bits = np.zeros((len(q), k), np.uint32)
    count = np.zeros(len(q), np.uint32)
    visits = np.zeros(len(q), np.uint32)
    d_full = np.zeros(len(q), np.uint32)
    for b in range(len(q)):
        r = oracle.scan(s, q[b], None, L, rescore, k)
        n = len(r["tid"])
        tid[b, :n] = r["tid"]
        dist_bits[b, :n] = r["dist"].view(np.uint32)
        count[b] = n
        visits[b] = r["stats"]["visits"]
        d_full[b] = r["stats"]["d_full"]
    return dict(tid=tid, dist_bits=dist_bits, count=count, visits=visits, d_full=d_full)


if __name__ == "__main__":
    out = {}
    for name in CASES:
        for key, val in run_case(name).items():
            out[f"{name}/{key}"] = val
    np.savez_compressed(os.path.join(HERE, "plain_golden.npz"), **out)
    print("wrote", len(out), "arrays")
