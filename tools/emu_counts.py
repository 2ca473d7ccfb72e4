"""Counts taken under the CPU SIMT emulator (tests/simt) - warp collectives per visit and per role for the measured
search kernel and for the HV=1 flavour with each alternative - as a proxy for the length of each warp's dependent chain
(NOT a timing).  Needs no GPU:   python tools/emu_counts.py > profiles/r01_emulation_counts.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")):
    sys.path.insert(0, p)
import emu  # noqa: E402
from conftest import build_case  # noqa: E402
from oracle import fixtures  # noqa: E402

s = build_case(6000, 768, 0, seed=5, kind="normal", R=50, L_build=100)
q = fixtures.gen_vectors(6, 768, 9, "normal")
_, qc = emu.prepare(s, q)
print("# 6000 x 768-d, 2-bit SBQ, R=50, oracle-built graph; 6 queries at L=150, 259 stream rows (bench operating point)")
print("# warp collectives per visit (a sub-warp collective counts by its share of the warp)")
rows = [("measured kernel (HV=0)", {}),
        ("HV=1, no alternative", {"DANN_HEAP_V2": 1, "DANN_HV_FLAGS": 0, "DANN_HV_NODE_ENTRIES": 0})]
names = {1: "register-path pushes", 2: "look-ahead pop", 4: "page-sized distance rounds", 8: "code-row prefetch",
         16: "nbr-row prefetch", 32: "32-ary visited search", 64: "no intra-list dedupe", 128: "root node by heap warp",
         256: "TID prefetch", 512: "fused expansion", 1024: "REDUX reductions", 2048: "uniform root prediction"}
for bit, nm in names.items():
    rows.append((f"HV=1 + {nm} ({bit})", {"DANN_HEAP_V2": 1, "DANN_HV_FLAGS": bit, "DANN_HV_NODE_ENTRIES": 0}))
rows.append(("HV=1, everything, node-carrying entries", {"DANN_HEAP_V2": 1}))
for name, env in rows:
    _, st, info = emu.search(s, qc, 150, 259, env=env, sm_count=1)
    v = sum(x["visits"] for x in st)
    print(f"{name:46s} controller {info['coll_even'] / v:6.1f}   heap warp {info['coll_odd'] / v:6.1f}   "
          f"(visits/query {v / len(st):.0f}, candidates/query {sum(x['candidates'] for x in st) / len(st):.0f})")
