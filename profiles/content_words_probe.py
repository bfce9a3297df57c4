#!/usr/bin/env python
"""BM25 top-10 of a batch of 64 over 1M docs (BASELINE configs[1] postings) for the two query
distributions of SURVEY 8d -- terms from the full Zipf law, and "content words" (ranks > 27 only:
real queries carry no stop-words) -- through both BM25 paths: the candidate-driven chain
(anr_bm25_ms.cu, default) and the exhaustive tiled scan (anr_bm25.cu, ANR_BM25_MAXSCORE=0), alone
and inside the hybrid step.  CUDA events; one JSON line.

    python profiles/content_words_probe.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    pkg = importlib.import_module("a-nice-rag_b200")
    env = bench.Env()
    env.torch, env.dist, env.rank, env.world = torch, None, 0, 1
    env.device = dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    env.engine, env.native = engine, native = pkg.engine, pkg.native
    env.synth = synth = importlib.import_module("a-nice-rag_b200.synth")
    env.sharded = importlib.import_module("a-nice-rag_b200.sharded")
    env.ctx = ctx = engine.context(0)
    vocab, B, K = 50_000, 64, bench.TOPK
    w = bench.Workload(env, 1_000_000, vocab, True)
    qb = bench.QueryBatch(env, B, vocab)
    nd = w.post["nd"].cpu().numpy().astype(np.int64)
    sets = {"zipf": qb.t_host,
            "content_words": synth.zipf_queries(B, bench.N_TERMS, vocab, bench.ZIPF_S,
                                                seed=bench.T_SEED + 1, skip_head=27)}
    k_sc = torch.empty((B, K), dtype=torch.float32, device=dev)
    k_id = torch.empty((B, K), dtype=torch.int32, device=dev)
    k_ct = torch.empty((B,), dtype=torch.int32, device=dev)

    def timed(fn, n=30):
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {"docs": 1_000_000, "batch": B, "unit": "ms per batch", "sets": {}}
    for name, terms in sets.items():
        t_dev = torch.from_numpy(terms.reshape(-1).copy()).to(dev)
        rec = {"postings_per_query": float(nd[terms.reshape(-1)].sum()) / B}

        def bm25_only():
            native.call("anr_bm25_search", ctx.handle, w.bm25.handle, t_dev.data_ptr(),
                        qb.off_dev.data_ptr(), B, K, None, None, 0, k_sc.data_ptr(), k_id.data_ptr(),
                        k_ct.data_ptr(), engine.torch_stream_ptr())

        def hybrid():
            native.call("anr_hybrid_search", ctx.handle, w.dense.handle, w.bm25.handle,
                        qb.q_dev.data_ptr(), t_dev.data_ptr(), qb.off_dev.data_ptr(), B, K, K, None,
                        None, None, 0, bench.W_DENSE, bench.W_BM25, bench.WRRF_K, K,
                        qb.out_ids.data_ptr(), qb.out_scores.data_ptr(), qb.out_counts.data_ptr(),
                        None, None, None, None, engine.torch_stream_ptr())
        got = {}
        for path, envval in (("candidate_driven", None), ("tiled_scan", "0")):
            if envval is None:
                os.environ.pop("ANR_BM25_MAXSCORE", None)
            else:
                os.environ["ANR_BM25_MAXSCORE"] = envval
            rec[path] = {"bm25_alone": timed(bm25_only), "hybrid_step": timed(hybrid)}
            bm25_only()
            torch.cuda.synchronize()
            got[path] = (k_id.cpu().numpy().copy(), k_sc.cpu().numpy().copy())
            if envval is None:
                rec[path]["rerun_queries"] = ctx.last_rerun()[1]
        os.environ.pop("ANR_BM25_MAXSCORE", None)
        a, b = got["candidate_driven"], got["tiled_scan"]
        rec["same_ids"] = float((a[0] == b[0]).mean())
        rec["max_score_diff"] = float(np.abs(a[1] - b[1]).max())
        out["sets"][name] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
