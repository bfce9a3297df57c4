"""BM25-only probe: builds a Zipf index on the device and runs anr_bm25_search (for sanitizer /
profiler runs).  Usage: python profiles/bm25_probe.py [n_docs] [vocab] [batch]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("a-nice-rag_b200")
engine, synth = pkg.engine, importlib.import_module("a-nice-rag_b200.synth")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
post = synth.zipf_postings_torch(n, vocab, 1.1, 2024, dev)
idf = synth.idf_from_counts(n, post["nd"].cpu().numpy(), 0.05)
avgdl = float(post["doc_len"].to(torch.int64).sum()) / n
index = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf,
                         1.7, 0.83, avgdl, n_terms=vocab, n_docs=n)
tq = synth.zipf_queries(b, 8, vocab, 1.1, seed=2025)
scores, docs, counts = index.search([list(map(int, t)) for t in tq], 10)
print("ok", n, index.n_postings, docs[0].tolist(), scores[0].tolist())
