"""Relation files of a checkpointed data directory -> the flat DANNSNP1 snapshot file (Snapshot.save_raw) that
harness/snapshot_raw.h, the sidecar and the C harnesses read.  Host only (dann_pg_*: pgvectorscale_b200/pgreader.py).

  python tools/pg_to_snapshot.py --index base/5/16402 --heap base/5/16390 --toast base/5/16393 --out snap.raw \\
         --dim 768 --R 50 --start 1:1 --means 120:1 [--dim-index 768 --bits 2 --distance 0 --atts 8d,-1i --plain]

The MetaPage scalars come from the caller (SELECT on the index's meta page through the extension, or its reloptions);
see INTEGRATION.md §4b for why the reader does not guess them."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _ip(s):
    b, o = s.split(":")
    return int(b), int(o)


def main(argv=None):
    from pgvectorscale_b200 import pgreader
    from pgvectorscale_b200.snapshot import default_bits
    ap = argparse.ArgumentParser()
    ap.add_argument("--index", required=True)
    ap.add_argument("--heap", required=True)
    ap.add_argument("--toast")
    ap.add_argument("--out", required=True)
    ap.add_argument("--dim", type=int, required=True)
    ap.add_argument("--dim-index", type=int)
    ap.add_argument("--bits", type=int)
    ap.add_argument("--R", type=int, required=True)
    ap.add_argument("--distance", type=int, default=0)
    ap.add_argument("--start", type=_ip, required=True)
    ap.add_argument("--means", type=_ip)
    ap.add_argument("--atts", default="", help="attlen+attalign of the columns in front of the vector column, e.g. 8d,-1i")
    ap.add_argument("--plain", action="store_true", help="storage_layout = plain")
    a = ap.parse_args(argv)
    dim_index = a.dim_index or a.dim
    meta = pgreader.PgMeta(num_dimensions=a.dim, num_dimensions_to_index=dim_index, bq_bits=a.bits or default_bits(dim_index),
                           num_neighbors=a.R, distance_type=a.distance, start=a.start, means=a.means)
    atts = [(int(t[:-1]), t[-1]) for t in a.atts.split(",") if t]
    with pgreader.PgRelation(a.index) as rel:
        info = rel.info()
        snap, _, fp, layout = rel.extract_plain(meta) if a.plain else rel.extract_sbq(meta)
    with pgreader.PgRelation(a.heap) as heap:
        toast = pgreader.PgRelation(a.toast) if a.toast else None
        try:
            snap.vectors, missing = pgreader.fetch_heap_vectors(heap, toast, snap.heap_tid, a.dim, atts)
        finally:
            if toast:
                toast.close()
    snap.save_raw(a.out)
    print(f"{snap.n} nodes, {missing} heap rows gone, root layout {layout}, relation fingerprint {fp:016x} "
          f"({info['nblocks']} blocks, max LSN {info['max_lsn']:x}) -> {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
