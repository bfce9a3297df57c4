"""Latency of the drop-in Python API (the calls a user of the reference makes), next to the
C-ABI numbers of bench.py: SearchEngine.similarity_search_with_embedding over an n-row DataFrame
(result = df.iloc[top].copy() + similarity column, like the reference) with and without the
"CG,NG" filter, and the three-call hybrid (dense + BM25 + WRRF) on a 20k corpus.
Usage: python profiles/api_latency.py [n_rows]"""
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("a-nice-rag_b200")
synth = importlib.import_module("a-nice-rag_b200.synth")
registry = pkg.registry
from oracle import bm25_okapi  # noqa: E402  (only builds the BM25Okapi-shaped fixture object)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = 1024
se = pkg.SearchEngine(None, None)


def p50(fn, iters=100, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return statistics.median(ts)


emb = synth.unit_vectors(n, d, seed=1234)
srcs = synth.sources(n, seed=77)
df = pd.DataFrame({"id": synth.chunk_ids(n, srcs), "document": [""] * n, "source": srcs,
                   "embedding": pd.Series(list(emb), dtype=object), "url": [""] * n})
registry.register_frame(df, emb).index()
queries = synth.unit_vectors(64, d, seed=4321)
it = iter(range(10**9))
out = {"rows": n, "d": d}
out["similarity_search_with_embedding_p50_ms"] = p50(
    lambda: se.similarity_search_with_embedding(queries[next(it) % 64], df, "voyage-3-large", 10))
out["similarity_search_with_embedding_filtered_p50_ms"] = p50(
    lambda: se.similarity_search_with_embedding(queries[next(it) % 64], df, "voyage-3-large", 10,
                                                "CG,NG"))

# hybrid through the three reference calls on the reference's own CPU-runnable size
n2, vocab = 20_000, 50_000
doc_ptr, tokens = synth.zipf_corpus(n2, vocab, 1.1, seed=2024)
okapi = bm25_okapi.BM25Okapi(synth.doc_token_lists(doc_ptr, tokens), k1=1.7, b=0.83, epsilon=0.05)
df2 = df.iloc[:n2].reset_index(drop=True)
df2.attrs.clear()
ids2 = df2["id"].tolist()
sections = [synth.Document("", {"id": ids2[i], "source": srcs[i]}) for i in range(n2)]
tq = synth.zipf_queries(64, 8, vocab, 1.1, seed=2025)
toks = [synth.token_strings(t) for t in tq]
weights = {"voyage-3-large": 5.0, "BM25": 1.0}


def hybrid():
    i = next(it) % 64
    res = se.similarity_search_with_embedding(queries[i], df2, "voyage-3-large", 10)
    hits = se.bm25_search_preprocessed(toks[i], okapi, sections, ids2, 10)
    return se.weighted_reciprocal_rank_fusion([(res["id"].tolist(), "voyage-3-large"),
                                               (hits, "BM25")], weights, 40)[:10]


hybrid()
out["hybrid_three_calls_20k_p50_ms"] = p50(hybrid)
out["hybrid_batch64_one_call_20k_ms"] = p50(
    lambda: se.hybrid_search_batch(queries, toks, df2, okapi, sections, ids2, weights,
                                   "voyage-3-large", 10, 10, 40), iters=20, warm=3)
print(json.dumps(out))
