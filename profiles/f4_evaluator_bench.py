#!/usr/bin/env python
"""SURVEY.md 8(f) f4 on the evaluator's shape (src/retrieval_eval.py:142-143, :276-281, :366-378):
thousands of queries x (dense model(s) + BM25), `similarity_k = common_sections_n = 12000`
(a full ranking of the filtered corpus), filter "CG,NG", wrrf_k = 40 -- on a synthetic 20k-chunk
corpus (the NICE corpus is ~10k CG/NG chunks).

  ours : a-nice-rag_b200.batch_retrieval.retrieve_documents_batch (one dense call per model, one
         BM25 call, one anr_wrrf_fuse over [B, lists, 12000] ids; lists stay on the device)
  cpu  : the reference loop -- oracle.orchestrator.retrieve_documents per query (the restatement
         pinned against the unmodified method) over oracle.cpu_search_engine (np.stack per call,
         literal BM25Okapi.get_scores, Python RRF) -- on a bounded sample of the same queries,
         timed on the box's host cores; the same queries are compared id for id

    python profiles/f4_evaluator_bench.py [--queries 2048] [--cpu-queries 6] [--models 1]
One JSON line.
"""
import argparse
import importlib
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=20_000)
    ap.add_argument("--queries", type=int, default=2048)
    ap.add_argument("--cpu-queries", type=int, default=6)
    ap.add_argument("--models", type=int, default=1, help="dense models (1..4) next to BM25")
    ap.add_argument("--k", type=int, default=12_000)
    args = ap.parse_args()
    import pandas as pd
    import torch
    from oracle import bm25_okapi, cpu_search_engine, orchestrator
    pkg = importlib.import_module("a-nice-rag_b200")
    batch = importlib.import_module("a-nice-rag_b200.batch_retrieval")
    synth = importlib.import_module("a-nice-rag_b200.synth")
    n, d, vocab = args.chunks, 1024, 50_000
    srcs = synth.sources(n, seed=7)
    ids = synth.chunk_ids(n, srcs)
    models = ["voyage-3-large", "voyage-3.5", "text-embedding-3-large", "Qwen3"][:args.models]
    frames = {}
    for mi, m in enumerate(models):
        emb = synth.unit_vectors(n, d, seed=1234 + mi)
        frames[m] = pd.DataFrame({"id": ids, "document": [f"doc {i}" for i in range(n)],
                                  "source": srcs, "embedding": list(emb), "url": [""] * n})
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=2024)
    okapi = bm25_okapi.BM25Okapi(synth.doc_token_lists(doc_ptr, tokens), k1=1.7, b=0.83, epsilon=0.05)
    sections = [types.SimpleNamespace(page_content=f"doc {i}", metadata={"id": ids[i], "source": srcs[i]})
                for i in range(n)]
    nice = pkg.InfoSource("nice")
    weights = {"voyage-3-large": 5.0, "voyage-3.5": 1.0, "text-embedding-3-large": 1.0, "Qwen3": 1.0,
               "BM25": 1.0}
    qe = {m: synth.unit_vectors(args.queries, d, seed=4321 + mi) for mi, m in enumerate(models)}
    tq = synth.zipf_queries(args.queries, 8, vocab, 1.1, seed=2025)
    toks = [synth.token_strings(t) for t in tq]
    kw = dict(similarity_k=args.k, common_sections_n=args.k, use_hybrid_search=True,
              filename_type_filter="CG,NG", model_weights=weights, wrrf_k=40)

    system = types.SimpleNamespace(config=pkg.Config(), search_engine=pkg.SearchEngine(None, None),
                                   embeddings_data={nice: frames}, bm25_data={nice: (okapi, sections, ids)})
    # warm-up: uploads the frames, inverts the BM25 object, builds masks / id codes
    t0 = time.perf_counter()
    batch.retrieve_documents_batch(system, {m: e[:64] for m, e in qe.items()}, None, toks[:64],
                                   info_source="NICE", **kw)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = batch.retrieve_documents_batch(system, qe, None, toks, info_source="NICE", **kw)
    torch.cuda.synchronize()
    ours_s = time.perf_counter() - t0

    cpu_system = types.SimpleNamespace(config=pkg.Config(),
                                       search_engine=cpu_search_engine.CpuSearchEngine(),
                                       embeddings_data={nice: frames},
                                       bm25_data={nice: (okapi, sections, ids)})
    sample = list(range(0, args.queries, max(1, args.queries // args.cpu_queries)))[:args.cpu_queries]
    t0 = time.perf_counter()
    want = [orchestrator.retrieve_documents(cpu_system, nice, {m: e[q] for m, e in qe.items()}, None,
                                            toks[q], **kw) for q in sample]
    cpu_s = time.perf_counter() - t0
    # id-for-id over the whole fused ranking; a mismatch may only be a swap inside a tie group
    # (equal fused scores), so the multisets must agree and the first difference is reported
    same = sum(got[q] == w for q, w in zip(sample, want))
    same_sets = sum(sorted(got[q]) == sorted(w) for q, w in zip(sample, want))
    first_diff = [next((i for i, (a, b) in enumerate(zip(got[q], w)) if a != b), None)
                  for q, w in zip(sample, want)]
    # how far apart are the two rankings?  A full ranking of ~13k chunks has pairs of dense scores
    # closer than fp32 summation order resolves (the reference's BLAS order is not ours) and BM25
    # scores that differ only beyond fp32: such pairs swap, and every swap moves the fused ranks
    # behind it by one -- displacements stay within a few positions
    moved, worst = [], []
    for q, w in zip(sample, want):
        pos = {doc: i for i, doc in enumerate(w)}
        disp = [abs(i - pos[doc]) for i, doc in enumerate(got[q]) if doc in pos]
        moved.append(sum(1 for x in disp if x))
        worst.append(max(disp) if disp else None)
    print(json.dumps({
        "workload": f"{n} chunks x {d}-d, {len(models)} dense model(s) + BM25 (V={vocab}), filter "
                    f"'CG,NG', similarity_k = common_sections_n = {args.k}, wrrf_k = 40 "
                    "(src/retrieval_eval.py:142-143, :276-281)",
        "queries": args.queries, "fused_list_len": len(got[0]),
        "ours": {"seconds": ours_s, "queries_per_s": args.queries / ours_s,
                 "first_call_setup_s": setup_s},
        "cpu_reference_loop": {"seconds": cpu_s, "queries": len(sample),
                               "queries_per_s": len(sample) / cpu_s, "cores": os.cpu_count(),
                               "kind": "port (orchestrator restatement + stock per-query search path)"},
        "speedup": (args.queries / ours_s) / (len(sample) / cpu_s),
        "parity": {"queries": len(sample), "identical_rankings": same,
                   "identical_id_sets": same_sets, "first_difference_at": first_diff,
                   "positions_that_moved": moved, "largest_rank_displacement": worst},
    }), flush=True)


if __name__ == "__main__":
    main()
