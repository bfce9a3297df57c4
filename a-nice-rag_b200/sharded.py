"""Corpus sharded by chunk across the GPUs of one box (one process per GPU).

Rank r owns a contiguous range of embedding rows AND the same documents' postings; queries
are replicated.  Every rank produces local top-k lists as sortable 64-bit keys (score in the
high word, GLOBAL id in the low word), the lists are exchanged with ONE NCCL all-gather per
retriever (``[B, k]`` keys per rank: latency-bound, a few hundred bytes at batch 1), every rank
merges them on the device and runs the weighted RRF on the merged lists -- fusion needs global
ranks, so it follows the merge.  BM25 shards score with GLOBAL statistics
(N, avgdl, idf), so a sharded search returns what the unsharded one does.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import engine, native


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank ``rank``: contiguous, sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_bm25_stats(nd_local, doc_len_sum_local: int, n_docs_local: int, group=None):
    """All-reduce of the document frequencies, token count and document count: every shard
    then derives the same idf table / avgdl as an unsharded index (run once at index build)."""
    import torch
    import torch.distributed as dist
    nd = nd_local.clone() if hasattr(nd_local, "clone") else torch.as_tensor(np.asarray(nd_local))
    extra = torch.tensor([doc_len_sum_local, n_docs_local], dtype=torch.int64, device=nd.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(nd, group=group)
        dist.all_reduce(extra, group=group)
    tokens, n_docs = int(extra[0]), int(extra[1])
    return nd, n_docs, tokens / max(n_docs, 1)


class ShardedHybrid:
    """Local indices of one rank + the exchange/merge/fusion of a sharded hybrid query."""

    def __init__(self, dense: engine.DenseIndex, bm25: engine.Bm25Index, row_base: int,
                 doc_base: Optional[int] = None, group=None):
        import torch
        self.torch = torch
        self.dense, self.bm25 = dense, bm25
        self.row_base = int(row_base)
        self.doc_base = int(row_base if doc_base is None else doc_base)
        self.group = group
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", engine.current_device())
        self._buf = {}

    def _buffers(self, b: int, k: int, top_n: int):
        key = (b, k, top_n)
        if key not in self._buf:
            t, dev = self.torch, self.device
            self._buf[key] = dict(
                local=t.empty((2, b, k), dtype=t.int64, device=dev),
                gathered=t.empty((self.world, 2, b, k), dtype=t.int64, device=dev),
                ids=t.empty((b, top_n), dtype=t.int32, device=dev),
                scores=t.empty((b, top_n), dtype=t.float64, device=dev),
                counts=t.empty((b,), dtype=t.int32, device=dev),
            )
        return self._buf[key]

    def phases(self, queries_dev, terms_dev, offsets_dev, b: int, k: int, w_dense: float,
               w_bm25: float, rrf_k: float, top_n: int):
        """The three parts of ``search`` as separate callables (local searches, all-gather,
        merge + fusion), for timing them one by one; each enqueues on torch's current stream."""
        ctx = engine.context(self.device.index)
        buf = self._buffers(b, k, top_n)
        local, gathered = buf["local"], buf["gathered"]

        def local_searches():
            native.call("anr_hybrid_search_keys", ctx.handle, self.dense.handle, self.bm25.handle,
                        queries_dev.data_ptr(), terms_dev.data_ptr(), offsets_dev.data_ptr(), b, k,
                        None, None, self.row_base, self.doc_base, local.data_ptr(),
                        engine.torch_stream_ptr())

        def exchange():
            if self.world > 1:
                self.dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)

        def merge_fuse():
            src = gathered if self.world > 1 else local
            native.call("anr_sharded_fuse", ctx.handle, src.data_ptr(), self.world, b, k,
                        float(w_dense), float(w_bm25), float(rrf_k), top_n, buf["ids"].data_ptr(),
                        buf["scores"].data_ptr(), buf["counts"].data_ptr(),
                        engine.torch_stream_ptr())
        return dict(local=local_searches, exchange=exchange, merge_fuse=merge_fuse)

    def search(self, queries_dev, terms_dev, offsets_dev, b: int, k: int, w_dense: float,
               w_bm25: float, rrf_k: float, top_n: int):
        """queries [b, d] fp32, CSR term ids int32: device tensors.  Returns device tensors
        (ids [b, top_n] int32 GLOBAL ids, scores f64, counts).  Enqueued on torch's current
        stream: 1 call for both local searches, 1 all-gather, 1 merge+fuse call."""
        t = self.torch
        ctx = engine.context(self.device.index)
        stream = engine.torch_stream_ptr()
        buf = self._buffers(b, k, top_n)
        local, gathered = buf["local"], buf["gathered"]
        # both local searches in one library call: the BM25 chain on the context's side stream beside
        # the dense chain, joined before the call's last kernel
        native.call("anr_hybrid_search_keys", ctx.handle, self.dense.handle, self.bm25.handle,
                    queries_dev.data_ptr(), terms_dev.data_ptr(), offsets_dev.data_ptr(), b, k, None,
                    None, self.row_base, self.doc_base, local.data_ptr(), stream)
        if self.world > 1:
            self.dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        else:
            gathered = local
        native.call("anr_sharded_fuse", ctx.handle, gathered.data_ptr(), self.world, b, k,
                    float(w_dense), float(w_bm25), float(rrf_k), top_n, buf["ids"].data_ptr(),
                    buf["scores"].data_ptr(), buf["counts"].data_ptr(), stream)
        return buf["ids"], buf["scores"], buf["counts"]


class ShardedHybridGraph:
    """One rank's sharded step of a FIXED shape -- local searches, NCCL all-gather, merge + fusion
    -- captured as ONE CUDA graph and replayed per batch ("keep in-stream,
    graph-captured").  A step on a small shard is a dozen short dependent launches around one
    tens-of-microseconds kernel; replay removes the launch gaps between them.  The capture owns a
    private context (the graph bakes in scratch addresses) and static input / output tensors;
    every rank must capture and replay in the same order (the collective is part of the graph)."""

    def __init__(self, dense: engine.DenseIndex, bm25: engine.Bm25Index, row_base: int, batch: int,
                 max_terms: int, k: int, w_dense: float, w_bm25: float, rrf_k: float, top_n: int,
                 doc_base: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = self.device = torch.device("cuda", dense.ctx_device)
        self.ctx = engine.Context(dense.ctx_device)
        self.dense, self.bm25 = dense, bm25
        self.batch, self.max_terms, self.k, self.top_n = int(batch), int(max_terms), int(k), int(top_n)
        self.row_base = int(row_base)
        self.doc_base = int(row_base if doc_base is None else doc_base)
        self.weights = (float(w_dense), float(w_bm25), float(rrf_k))
        i32, i64, f32, f64 = torch.int32, torch.int64, torch.float32, torch.float64
        self.q = torch.zeros((batch, dense.d), dtype=f32, device=dev)
        self.terms = torch.full((self.max_terms,), -1, dtype=i32, device=dev)
        self.offsets = torch.zeros((batch + 1,), dtype=i32, device=dev)
        self.local = torch.empty((2, batch, k), dtype=i64, device=dev)
        self.gathered = torch.empty((self.world, 2, batch, k), dtype=i64, device=dev)
        self.ids = torch.empty((batch, top_n), dtype=i32, device=dev)
        self.scores = torch.empty((batch, top_n), dtype=f64, device=dev)
        self.counts = torch.empty((batch,), dtype=i32, device=dev)
        self.stream = torch.cuda.Stream(dev)
        self.graph = None

    def _enqueue(self) -> None:
        stream = engine.torch_stream_ptr()
        native.call("anr_hybrid_search_keys", self.ctx.handle, self.dense.handle, self.bm25.handle,
                    self.q.data_ptr(), self.terms.data_ptr(), self.offsets.data_ptr(), self.batch,
                    self.k, None, None, self.row_base, self.doc_base, self.local.data_ptr(), stream)
        src = self.local
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered.view(-1), self.local.view(-1),
                                             group=self.group)
            src = self.gathered
        w_d, w_b, rrf_k = self.weights
        native.call("anr_sharded_fuse", self.ctx.handle, src.data_ptr(), self.world, self.batch,
                    self.k, w_d, w_b, rrf_k, self.top_n, self.ids.data_ptr(), self.scores.data_ptr(),
                    self.counts.data_ptr(), stream)

    def load(self, queries_dev, terms_dev, offsets_dev) -> None:
        n_terms = int(terms_dev.numel())
        if n_terms > self.max_terms:
            raise ValueError(f"{n_terms} query terms exceed the captured capacity {self.max_terms}")
        self.q.copy_(queries_dev, non_blocking=True)
        if n_terms:
            self.terms[:n_terms].copy_(terms_dev.reshape(-1), non_blocking=True)
        self.offsets.copy_(offsets_dev, non_blocking=True)

    def capture(self) -> None:
        """Warm-up (sizes the scratch, builds the lazily created index members, warms NCCL up),
        then the capture.  Collective: every rank calls it."""
        torch = self.torch
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for _ in range(3):
                self._enqueue()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
            self._enqueue()

    def replay(self):
        self.graph.replay()
        return self.ids, self.scores, self.counts
