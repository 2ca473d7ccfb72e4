"""SURVEY.md §8d config 1 AT ITS STATED SIZE: 100 000 x 768-d uniform[0,1) vectors, the reference's serial Vamana build
(oracle.build_graph: R = 50, L_build = 100, alpha = 1.2), one query, query_search_list_size = 100, query_rescore = 50,
k = 10; cosine with 2-bit SBQ (the reference's defaults at 768-d) and L2 with 1-bit SBQ.

The index itself is too large to commit (20 MB of neighbour lists per case) but it is a deterministic function of the
seeds, so what is committed (tests/golden/config1_golden.npz) is the oracle's ANSWER - row ids, rerank distance bits,
the consumed stream's node ids, the scan counters, a checksum of the neighbour lists - and the GPU test rebuilds the
index with the same serial build (about 100 s per case on one host core) before comparing.   python tests/golden/make_config1_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

N, DIM = 100_000, 768
SEED_DATA, SEED_QUERY = 0x5EED0001, 0x5EED0002
CASES = {"cosine_b2": (0, 2), "l2_b1": (1, 1)}       # name: (distance type, bits)
L, RESCORE, K = 100, 50, 10


def make_case(name):
    from oracle import fixtures
    dist, bits = CASES[name]
    v = fixtures.gen_vectors(N, DIM, SEED_DATA, "uniform")
    s = fixtures.make_index(v, dist, bits=bits, R=50, L_build=100, alpha=1.2)
    q = fixtures.gen_vectors(1, DIM, SEED_QUERY, "uniform")
    return s, q


def answer(s, q):
    from oracle import oracle
    r = oracle.scan(s, q[0], None, L, RESCORE, K, stream_cap=4096)
    return {"tid": r["tid"], "dist_bits": r["dist"].view(np.uint32), "stream": r["stream"],
            "counters": np.array([r["stats"][f] for f in ("visits", "d_quantized", "candidates", "d_full", "stream_len")], np.uint64),
            "nbrs_crc": np.array([zlib.crc32(np.ascontiguousarray(s.nbrs).tobytes())], np.uint64)}


if __name__ == "__main__":
    out = {}
    for name in CASES:
        s, q = make_case(name)
        for k, v in answer(s, q).items():
            out[f"{name}/{k}"] = v
        print(name, out[f"{name}/tid"][:4], out[f"{name}/counters"])
    np.savez_compressed(os.path.join(HERE, "config1_golden.npz"), **out)
