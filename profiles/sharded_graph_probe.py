#!/usr/bin/env python
"""The sharded hybrid step (anr_hybrid_search_keys -> NCCL all-gather -> anr_sharded_fuse) captured
as ONE CUDA graph per rank (a-nice-rag_b200/sharded.py::ShardedHybridGraph), and TWO batches in
flight on two captured steps / two NCCL groups, against the eager `ShardedHybrid.search`.  Every
variant's result is compared with the eager one before it is timed (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 profiles/sharded_graph_probe.py [--chunks 1000000] [--batch 64]

One JSON line on rank 0.  If capture of the NCCL collective fails on this torch / NCCL pair the
probe says so in the line instead of raising.
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--vocab", type=int, default=50_000)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    env = bench.Env()
    env.rank, local_rank, env.world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    env.torch, env.dist, env.device = torch, dist, torch.device("cuda", local_rank)
    if env.world > 1:
        dist.init_process_group("nccl", device_id=env.device)
    pkg = importlib.import_module("a-nice-rag_b200")
    env.engine, env.native = pkg.engine, pkg.native
    env.synth = importlib.import_module("a-nice-rag_b200.synth")
    env.sharded = importlib.import_module("a-nice-rag_b200.sharded")
    env.ctx = env.engine.context(local_rank)
    w = bench.Workload(env, args.chunks, args.vocab, True)
    qb = bench.QueryBatch(env, args.batch, args.vocab)
    B, K = args.batch, bench.TOPK
    weights = (bench.W_DENSE, bench.W_BM25, bench.WRRF_K)
    eager = env.sharded.ShardedHybrid(w.dense, w.bm25, row_base=w.lo)

    def step_eager():
        return eager.search(qb.q_dev, qb.t_dev, qb.off_dev, B, K, *weights, K)

    for _ in range(5):
        step_eager()
    torch.cuda.synchronize()
    want = [t.clone() for t in step_eager()]
    n = args.iters
    out = {"world": env.world, "chunks": args.chunks, "batch": B,
           "eager_ms": bench.timed(env, step_eager, n) / n}
    try:
        # one communicator per captured step: two NCCL kernels of ONE communicator must not run
        # concurrently, and the two-in-flight variant below overlaps the two steps
        groups = [None, dist.new_group() if env.world > 1 else None]
        caps = [env.sharded.ShardedHybridGraph(w.dense, w.bm25, w.lo, B, qb.t_dev.numel(), K, *weights,
                                               K, group=g) for g in groups]
        for c in caps:
            c.load(qb.q_dev, qb.t_dev, qb.off_dev)
            c.capture()
            c.replay()
        torch.cuda.synchronize()
        same = all(torch.equal(c.ids, want[0]) and torch.equal(c.scores, want[1]) and
                   torch.equal(c.counts, want[2]) for c in caps)
        out["graph_identical_to_eager"] = bool(same)
        out["graph_ms"] = bench.timed(env, caps[0].replay, n) / n

        def two_in_flight():
            cur = torch.cuda.current_stream()
            for c in caps:
                c.stream.wait_stream(cur)
                with torch.cuda.stream(c.stream):
                    c.graph.replay()
            for c in caps:
                cur.wait_stream(c.stream)
        two_in_flight()
        torch.cuda.synchronize()
        out["two_in_flight_ms_per_batch"] = bench.timed(env, two_in_flight, n // 2) / (n // 2) / 2
    except Exception as exc:   # report, do not raise: this is an experiment
        out["graph_error"] = repr(exc)[:400]
    if env.rank == 0:
        print(json.dumps(out), flush=True)
    # Captured graphs hold NCCL kernels of two communicators; tearing the process group down
    # with them alive hung the first run of this probe until its time limit.  Everything is
    # measured and printed: leave without running destructors.
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
