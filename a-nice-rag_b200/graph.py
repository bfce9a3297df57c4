"""A hybrid step of FIXED shape captured once as a CUDA graph and replayed per batch.

``anr_hybrid_search`` enqueues ~20 short kernels on two streams around its two long ones (the
dense pass and the BM25 scan).  With device pointers the call never synchronises and never reads
device results on the host (flagged tensor-core queries are rescanned on the device), so the
whole fork/join schedule can be captured: replay costs one launch instead of ~20 plus the event
traffic, which is what a latency-bound batch-1 query or a small shard pays for.

The capture owns a PRIVATE context: the graph bakes in addresses inside that context's scratch
buffer, which another call on a shared context could re-grow.  torch supplies the graph object
and the capture stream (plumbing); every node in the graph is a kernel or copy of this library.

SURVEY.md 8(d) asks for batch-1 latency "through CUDA-graph replay"; 8(e) for the exchange to
stay "in-stream, graph-captured".
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import engine, native


class HybridGraph:
    """dense top-k + BM25 top-k + weighted RRF for ``batch`` queries, replayable.

    Inputs are copied into static device buffers (``q`` [batch, d] fp32, ``terms`` [max_terms]
    int32 CSR term ids, ``offsets`` [batch + 1]); outputs are static device tensors
    (``ids`` [batch, top_n] int32, ``scores`` float64, ``counts`` int32 and, with
    ``want_lists``, the two per-retriever lists).  They are overwritten by the next replay.
    """

    def __init__(self, dense: engine.DenseIndex, bm25: engine.Bm25Index, batch: int,
                 max_terms: int, k_dense: int, k_bm25: int, w_dense: float, w_bm25: float,
                 rrf_k: float, top_n: int, row_mask: Optional[np.ndarray] = None,
                 doc_mask: Optional[np.ndarray] = None, doc_to_id=None, id_base: int = 0,
                 want_lists: bool = False):
        import torch
        if batch < 1 or max_terms < 1:
            raise ValueError("batch and max_terms must be >= 1")
        self.torch = torch
        self.dense, self.bm25 = dense, bm25
        self.batch, self.max_terms = int(batch), int(max_terms)
        dev = self.device = torch.device("cuda", dense.ctx_device)
        self.ctx = engine.Context(dense.ctx_device)      # private: see the module docstring
        i32, f32, f64 = torch.int32, torch.float32, torch.float64
        self.q = torch.zeros((batch, dense.d), dtype=f32, device=dev)
        self.terms = torch.full((self.max_terms,), -1, dtype=i32, device=dev)
        self.offsets = torch.zeros((batch + 1,), dtype=i32, device=dev)
        self.ids = torch.empty((batch, top_n), dtype=i32, device=dev)
        self.scores = torch.empty((batch, top_n), dtype=f64, device=dev)
        self.counts = torch.empty((batch,), dtype=i32, device=dev)
        self.lists = None
        if want_lists:
            self.lists = dict(dense_rows=torch.empty((batch, k_dense), dtype=i32, device=dev),
                              dense_scores=torch.empty((batch, k_dense), dtype=f32, device=dev),
                              bm25_ids=torch.empty((batch, k_bm25), dtype=i32, device=dev),
                              bm25_scores=torch.empty((batch, k_bm25), dtype=f32, device=dev))

        def mask_words(m):
            if m is None or hasattr(m, "data_ptr"):
                return m
            return torch.from_numpy(np.ascontiguousarray(m, dtype=np.uint32).view(np.int32)).to(dev)

        self._row_mask, self._doc_mask, self._doc_to_id = mask_words(row_mask), mask_words(doc_mask), doc_to_id
        lists = self.lists or {}
        self._args = (self.ctx.handle, dense.handle, bm25.handle, self.q.data_ptr(),
                      self.terms.data_ptr(), self.offsets.data_ptr(), self.batch, int(k_dense),
                      int(k_bm25), native.ptr(self._row_mask), native.ptr(self._doc_mask),
                      native.ptr(doc_to_id), int(id_base), float(w_dense), float(w_bm25),
                      float(rrf_k), int(top_n), self.ids.data_ptr(), self.scores.data_ptr(),
                      self.counts.data_ptr(), native.ptr(lists.get("dense_rows")),
                      native.ptr(lists.get("dense_scores")), native.ptr(lists.get("bm25_ids")),
                      native.ptr(lists.get("bm25_scores")))
        # Warm-up outside the capture: the scratch buffer reaches its size and the lazily built
        # index members (row-norm bound, bf16 shadow, BM25 head rows) exist -- those steps
        # synchronise, which a capture forbids.  Then the capture itself.
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._enqueue()
        side.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self._enqueue()

    def _enqueue(self) -> None:
        native.call("anr_hybrid_search", *self._args, engine.torch_stream_ptr())

    def load(self, queries, terms, offsets) -> None:
        """Copies one batch into the static input buffers on torch's current stream.  Accepts numpy
        arrays or torch tensors (host or device); ``offsets`` is the CSR row pointer of ``terms``."""
        t = self.torch
        as_t = lambda x: x if hasattr(x, "data_ptr") else t.from_numpy(np.ascontiguousarray(x))  # noqa: E731
        queries, terms, offsets = as_t(queries), as_t(terms).reshape(-1), as_t(offsets)
        if tuple(queries.shape) != tuple(self.q.shape):
            raise ValueError(f"queries must be {tuple(self.q.shape)}, got {tuple(queries.shape)}")
        if offsets.numel() != self.batch + 1:
            raise ValueError(f"offsets must hold {self.batch + 1} entries")
        n_terms = int(terms.numel())
        if n_terms > self.max_terms:
            raise ValueError(f"{n_terms} query terms exceed the captured capacity {self.max_terms}")
        self.q.copy_(queries, non_blocking=True)
        if n_terms:
            self.terms[:n_terms].copy_(terms, non_blocking=True)
        self.offsets.copy_(offsets, non_blocking=True)

    def replay(self):
        """One captured step on torch's current stream -> (ids, scores, counts) device tensors."""
        self.graph.replay()
        return self.ids, self.scores, self.counts

    def search(self, queries, terms, offsets):
        self.load(queries, terms, offsets)
        return self.replay()


class HybridPipeline:
    """Host buffers in, host buffers out, ``depth`` batches in flight.

    A synchronous call (``anr_hybrid_search`` with host pointers) returns when the results have
    landed, so the GPU idles while the host stages the next batch: at batch 64 over 1M chunks
    that is 0.69 ms per step against 0.63 ms of device time.  Here every slot owns a captured
    step (``HybridGraph``), a stream and pinned host buffers; ``submit`` enqueues H2D copy ->
    replay -> D2H copy on the slot's stream and returns at once, ``collect`` waits for the oldest
    slot only.  Results are those of the eager call, bit for bit, in submission order.
    (``tests/test_gpu_zz_pipeline.py``; ``bench.py`` reports it as ``e2e_pipelined``.)
    """

    def __init__(self, dense: engine.DenseIndex, bm25: engine.Bm25Index, batch: int, max_terms: int,
                 k_dense: int, k_bm25: int, w_dense: float, w_bm25: float, rrf_k: float,
                 top_n: int, depth: int = 2, **graph_kwargs):
        import torch
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.torch = torch
        self.batch, self.max_terms = int(batch), int(max_terms)
        dev = torch.device("cuda", dense.ctx_device)
        self.slots = []
        for _ in range(depth):
            g = HybridGraph(dense, bm25, batch, max_terms, k_dense, k_bm25, w_dense, w_bm25, rrf_k,
                            top_n, **graph_kwargs)
            pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()  # noqa: E731
            self.slots.append(dict(
                graph=g, stream=torch.cuda.Stream(dev), done=torch.cuda.Event(),
                q=pin(batch, dense.d, dtype=torch.float32), terms=pin(self.max_terms, dtype=torch.int32),
                offsets=pin(batch + 1, dtype=torch.int32), n_terms=0,
                ids=pin(batch, top_n, dtype=torch.int32), scores=pin(batch, top_n, dtype=torch.float64),
                counts=pin(batch, dtype=torch.int32), busy=False))
        self._next = 0          # slot the next submit uses
        self._oldest = 0        # slot the next collect waits for
        self._in_flight = 0

    def submit(self, queries: np.ndarray, terms: np.ndarray, offsets: np.ndarray) -> None:
        """Stages one batch (numpy arrays: [batch, d] fp32, CSR term ids int32, [batch + 1] offsets)
        and enqueues it.  Call ``collect`` first when ``depth`` batches are already in flight."""
        t = self.torch
        if self._in_flight == len(self.slots):
            raise RuntimeError("pipeline full: collect() the oldest batch first")
        s = self.slots[self._next]
        terms = np.ascontiguousarray(terms, dtype=np.int32).reshape(-1)
        if terms.shape[0] > self.max_terms:
            raise ValueError(f"{terms.shape[0]} query terms exceed the captured capacity {self.max_terms}")
        # host-side staging into the slot's pinned buffers (plain memcpy)
        s["q"].numpy()[...] = queries
        s["terms"].numpy()[:terms.shape[0]] = terms
        s["offsets"].numpy()[...] = offsets
        s["n_terms"] = int(terms.shape[0])
        g = s["graph"]
        with t.cuda.stream(s["stream"]):
            g.load(s["q"], s["terms"][:s["n_terms"]], s["offsets"])
            g.replay()
            s["ids"].copy_(g.ids, non_blocking=True)
            s["scores"].copy_(g.scores, non_blocking=True)
            s["counts"].copy_(g.counts, non_blocking=True)
            s["done"].record(s["stream"])
        s["busy"] = True
        self._next = (self._next + 1) % len(self.slots)
        self._in_flight += 1

    def collect(self):
        """Waits for the OLDEST batch in flight -> (ids [batch, top_n] int32, scores float64,
        counts int32) as numpy views of the slot's pinned buffers, valid until that slot is
        submitted to again (``depth`` submits later)."""
        if self._in_flight == 0:
            raise RuntimeError("nothing in flight")
        s = self.slots[self._oldest]
        s["done"].synchronize()
        s["busy"] = False
        self._oldest = (self._oldest + 1) % len(self.slots)
        self._in_flight -= 1
        return s["ids"].numpy(), s["scores"].numpy(), s["counts"].numpy()
