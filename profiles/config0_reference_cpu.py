#!/usr/bin/env python
"""BASELINE.json configs[0] through the UNMODIFIED reference on this container's CPU: 20k chunks x
1024-d fp32 + BM25 (V=50k, Zipf 1.1), 100 queries x 8 terms, top-10, weights 5:1, wrrf_k=40.
The reference's DatabaseManager loads a real SQLite chunks table and BM25 pickle, its SearchEngine
answers every query (similarity_search_with_embedding -> bm25_search_preprocessed -> weighted RRF);
the only stand-in is rank_bm25.BM25Okapi (absent offline) = oracle.bm25_okapi.BM25Okapi.
Runs only where /root/reference is mounted (it cannot travel to the GPU box).

    python profiles/config0_reference_cpu.py > profiles/r1_config0_reference_cpu.json
"""
import importlib
import json
import os
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import bm25_okapi, make_golden, reference_loader  # noqa: E402

synth = importlib.import_module("a-nice-rag_b200.synth")


def main():
    case = make_golden.config0_inputs()
    ref = reference_loader.load_reference()
    n = case["emb"].shape[0]
    srcs = list(case["sources"])
    ids = synth.chunk_ids(n, srcs)
    contents = [f"content {i}" for i in range(n)]
    bm25 = bm25_okapi.BM25Okapi(synth.doc_token_lists(case["doc_ptr"], case["tokens"]), k1=1.7, b=0.83,
                                epsilon=0.05)
    with tempfile.TemporaryDirectory() as tmp:
        db, pkl = os.path.join(tmp, "chunks.db"), os.path.join(tmp, "bm25.pkl")
        synth.write_chunks_db(db, ids, contents, srcs, case["emb"])
        synth.write_bm25_pickle(pkl, bm25, contents, ids, srcs)
        rank_mod = types.ModuleType("rank_bm25")
        rank_mod.BM25Okapi = bm25_okapi.BM25Okapi
        lc, lcs = types.ModuleType("langchain"), types.ModuleType("langchain.schema")
        lcd = types.ModuleType("langchain.schema.document")
        lcd.Document = synth.Document
        injected = {"rank_bm25": rank_mod, "langchain": lc, "langchain.schema": lcs,
                    "langchain.schema.document": lcd}
        saved = bm25_okapi.BM25Okapi.__module__
        sys.modules.update(injected)
        bm25_okapi.BM25Okapi.__module__ = "rank_bm25"
        try:
            dm = ref.DatabaseManager()
            t0 = time.perf_counter()
            df = dm.load_embeddings_from_sql(db, "voyage-3-large")
            r_bm25, r_sections, r_ids = dm.load_bm25_from_pickle(pkl)
            load_s = time.perf_counter() - t0
        finally:
            bm25_okapi.BM25Okapi.__module__ = saved
            for name in injected:
                sys.modules.pop(name, None)
    se = ref.SearchEngine(None, None)
    weights = {"voyage-3-large": 5.0, "BM25": 1.0}
    out = {}
    for flt in (None, "CG, NG"):
        t_dense = t_bm25 = t_fuse = 0.0
        nq = case["queries"].shape[0]
        for q in range(nq):
            toks = synth.token_strings(case["term_queries"][q])
            t0 = time.perf_counter()
            res = se.similarity_search_with_embedding(case["queries"][q], df, "voyage-3-large", 10, flt)
            t1 = time.perf_counter()
            hits = se.bm25_search_preprocessed(toks, r_bm25, r_sections, r_ids, 10, flt)
            t2 = time.perf_counter()
            fused = se.weighted_reciprocal_rank_fusion(
                [(res["id"].tolist(), "voyage-3-large"), (hits, "BM25")], weights, 40)[:10]
            t3 = time.perf_counter()
            assert len(fused) == 10
            t_dense += t1 - t0
            t_bm25 += t2 - t1
            t_fuse += t3 - t2
        total = t_dense + t_bm25 + t_fuse
        out["unfiltered" if flt is None else "filtered_CG_NG"] = {
            "dense_ms_per_query": 1e3 * t_dense / nq, "bm25_ms_per_query": 1e3 * t_bm25 / nq,
            "wrrf_ms_per_query": 1e3 * t_fuse / nq, "queries_per_s": nq / total}
    print(json.dumps({"config": "BASELINE configs[0]: 20000 chunks x 1024-d + BM25 V=50000, 100 queries "
                                "x 8 terms, top-10 WRRF (5:1, k=40)",
                      "impl": "unmodified reference SearchEngine / DatabaseManager (BM25Okapi restated)",
                      "host": {"cpus": os.cpu_count(), "numpy": np.__version__},
                      "load_s": load_s, **out}, indent=1))


if __name__ == "__main__":
    main()
