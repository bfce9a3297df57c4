#!/usr/bin/env python
"""Round-2 experiment -- NOT yet run on a GPU box.

The sharded hybrid step (anr_hybrid_search_keys -> NCCL all-gather -> anr_sharded_fuse) captured
as ONE CUDA graph per rank, and optionally TWO batches in flight on two private contexts /
streams, against the eager `ShardedHybrid.search` of bench.py.  Every variant's result is compared
with the eager one before it is timed (CUDA events, max over ranks).

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


class CapturedShard:
    """One rank's sharded step of a fixed shape, replayable.  Private context (the graph bakes in
    scratch addresses), static inputs / outputs, the all-gather inside the graph."""

    def __init__(self, torch, dist, engine, native, dense, bm25, row_base, world, b, n_terms, k,
                 weights, top_n, group=None):
        self.group = group
        self.torch, self.dist, self.engine, self.native = torch, dist, engine, native
        dev = torch.device("cuda", dense.ctx_device)
        self.ctx = engine.Context(dense.ctx_device)
        i32, i64, f32, f64 = torch.int32, torch.int64, torch.float32, torch.float64
        self.q = torch.zeros((b, dense.d), dtype=f32, device=dev)
        self.terms = torch.full((n_terms,), -1, dtype=i32, device=dev)
        self.offsets = torch.zeros((b + 1,), dtype=i32, device=dev)
        self.local = torch.empty((2, b, k), dtype=i64, device=dev)
        self.gathered = torch.empty((world, 2, b, k), dtype=i64, device=dev)
        self.ids = torch.empty((b, top_n), dtype=i32, device=dev)
        self.scores = torch.empty((b, top_n), dtype=f64, device=dev)
        self.counts = torch.empty((b,), dtype=i32, device=dev)
        self.dense, self.bm25 = dense, bm25
        self.args = (row_base, world, b, k, weights, top_n)
        self.stream = torch.cuda.Stream(dev)
        self.graph = None

    def enqueue(self):
        row_base, world, b, k, (w_d, w_b, rrf_k), top_n = self.args
        stream = self.engine.torch_stream_ptr()
        self.native.call("anr_hybrid_search_keys", self.ctx.handle, self.dense.handle,
                         self.bm25.handle, self.q.data_ptr(), self.terms.data_ptr(),
                         self.offsets.data_ptr(), b, k, None, None, row_base, row_base,
                         self.local.data_ptr(), stream)
        gathered = self.local
        if world > 1:
            self.dist.all_gather_into_tensor(self.gathered.view(-1), self.local.view(-1),
                                              group=self.group)
            gathered = self.gathered
        self.native.call("anr_sharded_fuse", self.ctx.handle, gathered.data_ptr(), world, b, k,
                         float(w_d), float(w_b), float(rrf_k), top_n, self.ids.data_ptr(),
                         self.scores.data_ptr(), self.counts.data_ptr(), stream)

    def load(self, q, terms, offsets):
        self.q.copy_(q, non_blocking=True)
        self.terms[:terms.numel()].copy_(terms, non_blocking=True)
        self.offsets.copy_(offsets, non_blocking=True)

    def capture(self):
        torch = self.torch
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(3):           # sizes the scratch, builds lazies, warms NCCL up
                self.enqueue()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
            self.enqueue()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local_rank, world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("a-nice-rag_b200")
    engine, native = pkg.engine, pkg.native
    synth = importlib.import_module("a-nice-rag_b200.synth")
    sharded = importlib.import_module("a-nice-rag_b200.sharded")
    lo, hi = sharded.shard_range(args.chunks, rank, world)
    emb, post = bench.make_shard(torch, dev, lo, hi, rank)
    nd, n_total, avgdl = sharded.global_bm25_stats(post["nd"], int(post["doc_len"].to(torch.int64).sum()),
                                                   hi - lo)
    idf = synth.idf_from_counts(n_total, nd.cpu().numpy(), bench.EPS)
    dense = engine.DenseIndex(emb, borrow=True)
    dense.set_shadow(True)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf,
                            bench.K1, bench.B_PARAM, avgdl, n_terms=bench.VOCAB, n_docs=hi - lo)
    B, K = args.batch, bench.TOPK
    q, terms, off = bench.make_queries(B)
    q_d = torch.from_numpy(q).to(dev)
    t_d = torch.from_numpy(terms.reshape(-1).copy()).to(dev)
    o_d = torch.from_numpy(off).to(dev)
    weights = (bench.W_DENSE, bench.W_BM25, bench.WRRF_K)
    eager = sharded.ShardedHybrid(dense, bm25, row_base=lo)

    def step_eager():
        return eager.search(q_d, t_d, o_d, B, K, *weights, K)

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(5):
        step_eager()
    torch.cuda.synchronize()
    want = [t.clone() for t in step_eager()]
    out = {"world": world, "chunks": args.chunks, "batch": B, "eager_ms": timed(step_eager, args.iters)}
    try:
        # one communicator per captured step: two NCCL kernels of ONE communicator must not run
        # concurrently, and the two-in-flight variant below overlaps the two steps
        groups = [None, dist.new_group() if world > 1 else None]
        caps = [CapturedShard(torch, dist, engine, native, dense, bm25, lo, world, B, t_d.numel(), K,
                              weights, K, group=g) for g in groups]
        for c in caps:
            c.load(q_d, t_d, o_d)
            c.capture()
            c.graph.replay()
        torch.cuda.synchronize()
        same = all(torch.equal(c.ids, want[0]) and torch.equal(c.scores, want[1]) and
                   torch.equal(c.counts, want[2]) for c in caps)
        out["graph_identical_to_eager"] = bool(same)
        out["graph_ms"] = timed(caps[0].graph.replay, args.iters)

        # two batches in flight: alternate the two captured steps on their own streams (every rank
        # issues them in the same order, so the NCCL calls pair up)
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
        out["two_in_flight_ms_per_batch"] = timed(two_in_flight, args.iters // 2) / 2
    except Exception as exc:   # report, do not raise: this is an experiment
        out["graph_error"] = repr(exc)[:400]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
