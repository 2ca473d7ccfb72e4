"""Regenerates tests/golden/scan_golden.npz from the CPU oracle.

The reference holds no golden id vectors for this path and cannot be run
here (Rust + Postgres), so these vectors are SELF-generated: they freeze the oracle's answers
(after it was pinned against the reference's KATs and the independent Python restatement) so
that later edits to oracle.cpp cannot drift silently, and they give the GPU tests inputs with
known answers.   python tests/golden/make_golden.py
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
    # name: (n, dim, distance, bits, labels, L, rescore, k, kind)
    "cos768b2": (600, 768, 0, 2, False, 100, 50, 10, "normal"),
    "l2_768b1": (600, 768, 1, 1, False, 50, 20, 10, "uniform"),
    "ip100lab": (500, 100, 2, 2, True, 30, 0, 12, "uniform"),
}


def make_case(name):
    from conftest import build_case
    from oracle import fixtures
    n, dim, dist, bits, labels, L, rescore, k, kind = CASES[name]
    s = build_case(n, dim, dist, bits=bits, seed=101, kind=kind, R=32, L_build=64, labels=labels,
                   deleted_every=11)
    q = fixtures.gen_vectors(8, dim, 202, kind)
    lab = [[1 + (i % 16)] for i in range(8)] if labels else None
    return s, q, lab, L, rescore, k


def run_case(name):
    from oracle import oracle
    s, q, lab, L, rescore, k = make_case(name)
    off = vals = None
    if lab is not None:
        off = np.zeros(len(lab) + 1, np.int32)
        vals = []
        for i, ls in enumerate(lab):
            vals.extend(ls)
            off[i + 1] = len(vals)
        vals = np.asarray(vals, np.int16)
    tid, dist, count, stats = oracle.scan_batch(s, q, vals, off, L, rescore, k, threads=1)
    return dict(tid=tid, dist_bits=dist.view(np.uint32), count=count,
                visits=stats["visits"], d_quantized=stats["d_quantized"])


if __name__ == "__main__":
    out = {}
    for name in CASES:
        for key, val in run_case(name).items():
            out[f"{name}/{key}"] = val
    np.savez_compressed(os.path.join(HERE, "scan_golden.npz"), **out)
    print("wrote", len(out), "arrays")
