#!/usr/bin/env python
"""Eager C-ABI call vs CUDA-graph replay of the same hybrid step (BASELINE configs[1] corpus:
1M chunks x 1024-d + BM25 over 1M docs), device-resident inputs, CUDA events on the launching
stream.  Prints one JSON line; every replayed result is compared with the eager one first.

    python profiles/graph_probe.py [--chunks 1000000] [--iters 100]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("a-nice-rag_b200")
    engine, native, synth = pkg.engine, pkg.native, importlib.import_module("a-nice-rag_b200.synth")
    graph = importlib.import_module("a-nice-rag_b200.graph")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    emb, post = bench.make_shard(torch, dev, 0, args.chunks, 0)
    idf = synth.idf_from_counts(args.chunks, post["nd"].cpu().numpy(), bench.EPS)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / args.chunks
    dense = engine.DenseIndex(emb, borrow=True)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf,
                            bench.K1, bench.B_PARAM, avgdl, n_terms=bench.VOCAB, n_docs=args.chunks)
    ctx = engine.context(0)
    K = bench.TOPK
    out = {"chunks": args.chunks, "iters": args.iters, "cases": []}

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for batch, shadow in ((1, True), (1, False), (64, True)):
        dense.set_shadow(shadow)
        q, terms, off = bench.make_queries(batch)
        q_d = torch.from_numpy(q).to(dev)
        t_d = torch.from_numpy(terms.reshape(-1).copy()).to(dev)
        o_d = torch.from_numpy(off).to(dev)
        ids = torch.empty((batch, K), dtype=torch.int32, device=dev)
        sc = torch.empty((batch, K), dtype=torch.float64, device=dev)
        ct = torch.empty((batch,), dtype=torch.int32, device=dev)

        def eager():
            native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle, q_d.data_ptr(),
                        t_d.data_ptr(), o_d.data_ptr(), batch, K, K, None, None, None, 0,
                        bench.W_DENSE, bench.W_BM25, bench.WRRF_K, K, ids.data_ptr(), sc.data_ptr(),
                        ct.data_ptr(), None, None, None, None, engine.torch_stream_ptr())

        for _ in range(5):
            eager()
        g = graph.HybridGraph(dense, bm25, batch, terms.size, K, K, bench.W_DENSE, bench.W_BM25,
                              bench.WRRF_K, K)
        g.load(q_d, t_d, o_d)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(g.ids, ids) and torch.equal(g.scores, sc) and torch.equal(g.counts, ct))
        ms_eager = timed(eager, args.iters)
        ms_graph = timed(g.replay, args.iters)
        ms_graph_load = timed(lambda: g.search(q_d, t_d, o_d), args.iters)
        out["cases"].append({"batch": batch, "dense": "bf16 shadow GEMM pass" if shadow else
                             "fp32 CUDA-core scan", "identical_results": same,
                             "eager_ms": ms_eager, "graph_replay_ms": ms_graph,
                             "graph_load_and_replay_ms": ms_graph_load})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
