#!/usr/bin/env python
"""SURVEY 8(f) f1 on the CPU: DatabaseManager.load_embeddings_from_sql of this repo (bulk BLOB
reader in the library + text columns fetched concurrently) against the UNMODIFIED reference loader
(src/database_manager.py:22-75, imported from /root/reference -- so this runs only where the
reference is mounted).  No GPU involved: without a device the product loader stops before the
upload, which is what is timed on both sides.

    python profiles/loader_bench.py [--rows 100000] [--dim 1024] [--content-bytes 12]
"""
import argparse
import importlib
import json
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--content-bytes", type=int, default=12,
                    help="length of the `content` text per row (NICE chunks are ~2-3 KB)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    logging.disable(logging.CRITICAL)
    synth = importlib.import_module("a-nice-rag_b200.synth")
    pkg = importlib.import_module("a-nice-rag_b200")
    from oracle import reference_loader
    n, d = args.rows, args.dim
    emb = synth.unit_vectors(n, d, seed=1)
    srcs = synth.sources(n, seed=2)
    ids = synth.chunk_ids(n, srcs)
    pad = "x" * max(args.content_bytes - 8, 0)
    with tempfile.TemporaryDirectory() as tmp:
        db = os.path.join(tmp, "chunks.db")
        synth.write_chunks_db(db, ids, [f"{i:08d}{pad}" for i in range(n)], srcs, emb)
        ref = reference_loader.load_reference()
        pkg.DatabaseManager().load_embeddings_from_sql(db, "warm")       # imports, page cache
        ours, theirs = [], []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            df = pkg.DatabaseManager().load_embeddings_from_sql(db, "m")
            ours.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            rdf = ref.DatabaseManager().load_embeddings_from_sql(db, "m")
            theirs.append(time.perf_counter() - t0)
        assert df["id"].tolist() == rdf["id"].tolist() and df["document"].tolist() == rdf["document"].tolist()
        assert all(np.array_equal(a, b) for a, b in zip(df["embedding"].values[::997], rdf["embedding"].values[::997]))
        print(json.dumps({"rows": n, "dim": d, "content_bytes": args.content_bytes,
                          "db_mib": os.path.getsize(db) >> 20, "cpus": os.cpu_count(),
                          "product_s": [round(x, 3) for x in ours], "reference_s": [round(x, 3) for x in theirs],
                          "speedup_best": round(min(theirs) / min(ours), 2)}))


if __name__ == "__main__":
    main()
