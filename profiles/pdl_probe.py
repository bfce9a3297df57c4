#!/usr/bin/env python
"""Which programmatic launches pay?  The hybrid step of BASELINE configs[1] (1M chunks x 1024-d +
BM25 over 1M docs, batch 64, device-resident inputs, eager C-ABI calls) under every setting of the
"pdl" option (include/anr_b200.h: bit 0 dense chain, bit 1 BM25 chain, bit 2 dense main kernel),
interleaved round by round so that the box's power-cap drift hits every setting alike; plus the
batch-1 step.  CUDA events on the launching stream; prints one JSON line.

    python profiles/pdl_probe.py [--rounds 5] [--steps 30]
"""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=1_000_000)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--masks", default="0,1,2,3,5,7")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("a-nice-rag_b200")
    env = bench.Env()
    env.torch, env.dist, env.rank, env.world = torch, None, 0, 1
    env.device = dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    env.engine, env.native = engine, native = pkg.engine, pkg.native
    env.synth = importlib.import_module("a-nice-rag_b200.synth")
    env.sharded = importlib.import_module("a-nice-rag_b200.sharded")
    env.ctx = ctx = engine.context(0)
    w = bench.Workload(env, args.chunks, bench.VOCAB if hasattr(bench, "VOCAB") else 50_000, True)
    dense, bm25 = w.dense, w.bm25
    K = bench.TOPK
    masks = [int(m) for m in args.masks.split(",")]

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {"chunks": args.chunks, "rounds": args.rounds, "steps": args.steps, "unit": "ms per step",
           "pdl_bits": "1 dense chain, 2 BM25 chain, 4 dense main kernel", "batches": {}}
    for batch in (64, 1):
        qb = bench.QueryBatch(env, batch, w.vocab)
        q_d, t_d, o_d = qb.q_dev, qb.t_dev, qb.off_dev
        ids, sc, ct = qb.out_ids, qb.out_scores, qb.out_counts

        def step():
            native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle, q_d.data_ptr(),
                        t_d.data_ptr(), o_d.data_ptr(), batch, K, K, None, None, None, 0,
                        bench.W_DENSE, bench.W_BM25, bench.WRRF_K, K, ids.data_ptr(), sc.data_ptr(),
                        ct.data_ptr(), None, None, None, None, engine.torch_stream_ptr())

        native.call("anr_set_option", b"pdl", 0)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        want = (ids.clone(), sc.clone(), ct.clone())
        ms = {m: [] for m in masks}
        same = {}
        for _ in range(args.rounds):
            for m in masks:
                native.call("anr_set_option", b"pdl", m)
                for _ in range(3):
                    step()
                ms[m].append(timed(step, args.steps))
                same[m] = bool(torch.equal(ids, want[0]) and torch.equal(sc, want[1])
                               and torch.equal(ct, want[2]))
        out["batches"][str(batch)] = {str(m): {"median": statistics.median(v), "all": v,
                                               "identical_to_serialised": same[m]}
                                      for m, v in ms.items()}
    native.call("anr_set_option", b"pdl", 7)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
