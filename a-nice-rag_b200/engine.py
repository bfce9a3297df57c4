"""Device-resident indices and the calls on them (thin Python over the C ABI).

Host-side work kept here, as SURVEY.md 8(b) assigns it: token string -> term id
dictionary, doc index <-> id maps, the CSR inversion of a ``BM25Okapi``-shaped
object, filter-string -> bit mask.  All scoring / selection / fusion runs in the
CUDA library; nothing in this module computes a score on the CPU.
"""
from __future__ import annotations

import ctypes as C
import re
import threading
from itertools import chain
from operator import methodcaller
import weakref
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import native


# ---------------------------------------------------------------------------------
# contexts: one per (thread, device) -- a context owns scratch memory and a stream and
# must not be shared by two threads (include/anr_b200.h)
# ---------------------------------------------------------------------------------
class Context:
    def __init__(self, device: int = 0):
        self.device = device
        handle = C.c_void_p()
        native.call("anr_ctx_create", device, C.byref(handle))
        self.handle = handle
        self._finalizer = weakref.finalize(self, native.load().anr_ctx_destroy, handle)

    def sync(self) -> None:
        native.call("anr_ctx_sync", self.handle)

    def last_rerun(self) -> Tuple[int, int]:
        """(dense queries, BM25 queries) of the last search call on this context that left the fast
        paths and were rerun exactly on the device; -1 = the call had no such path (anr_b200.h)."""
        d, b = C.c_int32(), C.c_int32()
        native.call("anr_ctx_last_rerun", self.handle, C.byref(d), C.byref(b))
        return d.value, b.value

    def info(self) -> Tuple[int, int, int]:
        sm, total, free = C.c_int32(), C.c_int64(), C.c_int64()
        native.call("anr_ctx_info", self.handle, C.byref(sm), C.byref(total), C.byref(free))
        return sm.value, total.value, free.value


_tls = threading.local()


def current_device() -> int:
    """The process' CUDA device: LOCAL_RANK-style selection goes through torch if it is loaded."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:  # pragma: no cover - torch is plumbing only
        pass
    return 0


def torch_stream_ptr() -> int:
    """torch's current stream as the ABI's ``stream`` argument.  torch's default stream is the
    legacy default stream, whose handle is 0 -- which the ABI reads as "use the context's own
    stream" -- so it is passed as cudaStreamLegacy (0x1) instead."""
    import torch
    return torch.cuda.current_stream().cuda_stream or 1


def _sync_producer_stream(*tensors) -> None:
    """Device tensors handed to the library may still be being written on torch's current stream;
    the library reads them on its own (non-blocking) stream, so wait for the producer once."""
    for x in tensors:
        if hasattr(x, "is_cuda") and x.is_cuda:
            import torch
            torch.cuda.current_stream(x.device).synchronize()
            return


def context(device: Optional[int] = None) -> Context:
    device = current_device() if device is None else device
    cache = getattr(_tls, "contexts", None)
    if cache is None:
        cache = _tls.contexts = {}
    ctx = cache.get(device)
    if ctx is None:
        ctx = cache[device] = Context(device)
    return ctx


# ---------------------------------------------------------------------------------
# filters
# ---------------------------------------------------------------------------------
def parse_prefixes(filename_type_filter: str) -> Tuple[str, ...]:
    """Same normalisation as src/search_engine.py:39 and :222."""
    return tuple(p.strip().upper() for p in filename_type_filter.split(","))


def prefix_mask(sources: Sequence[Optional[str]], filename_type_filter: str,
                frame_semantics: bool = False) -> np.ndarray:
    """bool[n]: upper-cased source starts with any prefix (None / non-string -> False).

    ``frame_semantics``: the DataFrame filter (search_engine.py:41-46) joins SEVERAL prefixes
    into the un-escaped regex ``^(?:A|B)``, so a prefix holding a regex metacharacter acts as a
    pattern there (and a malformed one raises, which the search methods turn into an empty
    result); a single prefix, and the BM25 filter (:224-231) always, compare literally.
    An all-string column goes through pandas' native string kernels (0.2 s per million sources
    instead of 0.5 s for the Python loop, which remains for columns holding non-strings) -- a
    cost paid once per filter string, by the first query that uses it.
    """
    import pandas as pd
    prefixes = parse_prefixes(filename_type_filter)
    n = len(sources)
    if n == 0:
        return np.zeros(0, dtype=bool)
    pattern = None
    if frame_semantics and len(prefixes) > 1 and any(re.escape(p) != p for p in prefixes):
        pattern = re.compile("^(?:" + "|".join(prefixes) + ")")   # a malformed one raises here
    # dtype inferred: an all-string (or string / None) column becomes pandas' native string array
    # with C++ kernels; anything else stays object and is walked in Python like before
    column = pd.Series(np.asarray(sources, dtype=object))
    if column.dtype != object:
        upper = column.str.upper()
        hit = (upper.str.startswith(prefixes, na=False) if pattern is None
               else upper.str.contains(pattern.pattern, na=False, regex=True))
        return hit.to_numpy(dtype=bool)
    out = np.zeros(n, dtype=bool)
    for i, src in enumerate(sources):
        if isinstance(src, str):
            up = src.upper()
            out[i] = up.startswith(prefixes) if pattern is None else pattern.search(up) is not None
    return out


def pack_mask(mask: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32 words, bit (i & 31) of word (i >> 5) = mask[i] (the ABI's layout)."""
    n = int(mask.shape[0])
    words = (n + 31) // 32
    padded = np.zeros(words * 32, dtype=np.uint8)
    padded[:n] = mask
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view("<u4").copy()


def _as_f32_matrix(x) -> np.ndarray:
    a = np.asarray(x)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------
# dense index
# ---------------------------------------------------------------------------------
class DenseIndex:
    """Row-major [n, d] fp32 matrix resident in HBM (the chunk-embedding matrix)."""

    def __init__(self, embeddings=None, n: Optional[int] = None, d: Optional[int] = None,
                 device: Optional[int] = None, borrow: bool = False):
        self.ctx_device = current_device() if device is None else device
        ctx = context(self.ctx_device)
        self._keepalive = None
        if embeddings is not None:
            if hasattr(embeddings, "data_ptr"):      # torch tensor, host or device
                if embeddings.dim() != 2 or str(embeddings.dtype) != "torch.float32":
                    raise ValueError("embeddings tensor must be 2-D float32")
                embeddings = embeddings.contiguous()
                n, d = int(embeddings.shape[0]), int(embeddings.shape[1])
                if borrow:
                    self._keepalive = embeddings
            else:
                embeddings = _as_f32_matrix(embeddings)
                n, d = embeddings.shape
                borrow = False
        elif n is None or d is None:
            raise ValueError("give embeddings or (n, d)")
        handle = C.c_void_p()
        _sync_producer_stream(embeddings)
        native.call("anr_dense_create", ctx.handle, native.ptr(embeddings), int(n), int(d),
                    1 if borrow else 0, C.byref(handle))
        self.handle = handle
        self.n, self.d = int(n), int(d)
        self._masks: Dict[str, object] = {}
        self._finalizer = weakref.finalize(self, native.load().anr_dense_destroy, handle)

    def upload(self, row0: int, rows) -> None:
        rows = rows if hasattr(rows, "data_ptr") else _as_f32_matrix(rows)
        _sync_producer_stream(rows)
        native.call("anr_dense_upload", context(self.ctx_device).handle, self.handle, int(row0),
                    native.ptr(rows), int(rows.shape[0]))

    def set_shadow(self, enable: bool = True) -> None:
        """Batches of more than 32 queries nominate candidates on a bf16 copy of the matrix
        (half the HBM bytes per pass); results are unchanged (exact fp32 rescoring)."""
        native.call("anr_dense_set_shadow", self.handle, 1 if enable else 0)

    def invalidate(self) -> None:
        """The rows were rewritten in place by their owner (a borrowed tensor): drop the cached
        row-norm bound and the bf16 shadow copy; the next search rebuilds them."""
        native.call("anr_dense_invalidate", self.handle)

    def search(self, queries, k: int, row_mask: Optional[np.ndarray] = None, id_base: int = 0):
        """-> (scores f32 [b, k], rows i32 [b, k], counts i32 [b]); rows = -1 past counts."""
        q = _as_f32_matrix(queries)
        if q.shape[1] != self.d:
            raise ValueError(f"query width {q.shape[1]} != index width {self.d}")
        b = q.shape[0]
        scores = np.empty((b, k), dtype=np.float32)
        rows = np.empty((b, k), dtype=np.int32)
        counts = np.empty(b, dtype=np.int32)
        native.call("anr_dense_search", context(self.ctx_device).handle, self.handle, native.ptr(q),
                    b, int(k), native.ptr(row_mask), int(id_base), native.ptr(scores),
                    native.ptr(rows), native.ptr(counts), None)
        return scores, rows, counts


    def search_device(self, queries_dev, k: int, row_mask=None, id_base: int = 0):
        """Same search with DEVICE tensors in and out (torch is the allocator): queries [b, d] fp32
        on this index' GPU -> (scores f32 [b, k], rows i32 [b, k], counts i32 [b]) device tensors,
        enqueued on torch's current stream without synchronising (results stay in HBM for the next
        kernel, e.g. anr_wrrf_fuse)."""
        import torch
        b = int(queries_dev.shape[0])
        dev = queries_dev.device
        scores = torch.empty((b, k), dtype=torch.float32, device=dev)
        rows = torch.empty((b, k), dtype=torch.int32, device=dev)
        counts = torch.empty((b,), dtype=torch.int32, device=dev)
        native.call("anr_dense_search", context(self.ctx_device).handle, self.handle,
                    queries_dev.data_ptr(), b, int(k), native.ptr(row_mask), int(id_base),
                    scores.data_ptr(), rows.data_ptr(), counts.data_ptr(), torch_stream_ptr())
        return scores, rows, counts


# ---------------------------------------------------------------------------------
# BM25 index
# ---------------------------------------------------------------------------------
def invert_okapi(bm25) -> Tuple[Dict[str, int], np.ndarray, np.ndarray, np.ndarray, np.ndarray,
                                np.ndarray]:
    """CSR inversion of a BM25Okapi-shaped object (attributes of rank_bm25.BM25Okapi).

    Returns (vocab, term_ptr i64 [V+1], post_doc i32, post_tf i32, doc_len i32, idf f64 [V]).
    Term ids follow the order of ``bm25.idf`` (the package's vocabulary order); postings of a
    term are in ascending document order because documents are visited in order.
    """
    vocab = {tok: i for i, tok in enumerate(bm25.idf.keys())}
    n_terms = len(vocab)
    doc_freqs = bm25.doc_freqs
    n_docs = len(doc_freqs)
    doc_sizes = np.fromiter(map(len, doc_freqs), dtype=np.int64, count=n_docs)
    total = int(doc_sizes.sum())
    # one pass over all (document, term) pairs in document order: no per-document arrays
    terms = np.fromiter(map(vocab.__getitem__, chain.from_iterable(doc_freqs)), dtype=np.int32,
                        count=total)
    tfs = np.fromiter(chain.from_iterable(map(methodcaller("values"), doc_freqs)), dtype=np.int32,
                      count=total)
    doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(doc_sizes, out=doc_ptr[1:])
    idf = np.fromiter((bm25.idf[t] for t in vocab), dtype=np.float64, count=n_terms)
    doc_len = np.asarray(bm25.doc_len, dtype=np.int32)
    try:   # documents x terms (CSR) -> terms x documents: a linear-time, order-preserving transpose
        from scipy import sparse
        # (csr -> csc is a counting sort by column that keeps the row order inside a column and
        # neither sorts nor merges entries; a dict cannot repeat a term anyway)
        by_term = sparse.csr_matrix((tfs, terms, doc_ptr), shape=(n_docs, max(n_terms, 1))).tocsc()
        term_ptr = by_term.indptr.astype(np.int64)[:n_terms + 1]
        if n_terms == 0:
            term_ptr = np.zeros(1, dtype=np.int64)
        return (vocab, term_ptr, by_term.indices.astype(np.int32), by_term.data.astype(np.int32),
                doc_len, idf)
    except ImportError:
        pass
    docs = np.repeat(np.arange(n_docs, dtype=np.int32), doc_sizes)
    order = np.argsort(terms, kind="stable")      # stable: documents stay ascending per term
    term_ptr = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(np.bincount(terms, minlength=n_terms), out=term_ptr[1:])
    return vocab, term_ptr, docs[order], tfs[order], doc_len, idf


class Bm25Index:
    """CSR inverted index in HBM with the BM25Okapi posting weights precomputed."""

    def __init__(self, term_ptr, post_doc, post_tf, doc_len, idf, k1: float, b: float,
                 avgdl: float, vocab: Optional[Dict[str, int]] = None,
                 device: Optional[int] = None, n_terms: Optional[int] = None,
                 n_docs: Optional[int] = None):
        self.ctx_device = current_device() if device is None else device
        ctx = context(self.ctx_device)

        def prep(x, dtype):
            if hasattr(x, "data_ptr"):
                return x.contiguous()
            return np.ascontiguousarray(x, dtype=dtype)

        term_ptr, post_doc, post_tf = prep(term_ptr, np.int64), prep(post_doc, np.int32), prep(post_tf, np.int32)
        doc_len, idf = prep(doc_len, np.int32), prep(idf, np.float64)
        self.n_terms = int(term_ptr.shape[0]) - 1 if n_terms is None else int(n_terms)
        self.n_docs = int(doc_len.shape[0]) if n_docs is None else int(n_docs)
        handle = C.c_void_p()
        _sync_producer_stream(term_ptr, post_doc, post_tf, doc_len, idf)
        native.call("anr_bm25_create", ctx.handle, native.ptr(term_ptr), native.ptr(post_doc),
                    native.ptr(post_tf), native.ptr(doc_len), native.ptr(idf), self.n_terms,
                    self.n_docs, float(k1), float(b), float(avgdl), C.byref(handle))
        self.handle = handle
        self.vocab = vocab
        nnz = C.c_int64()
        native.call("anr_bm25_shape", handle, None, None, C.byref(nnz))
        self.n_postings = nnz.value
        self._masks: Dict[str, object] = {}
        self._finalizer = weakref.finalize(self, native.load().anr_bm25_destroy, handle)

    @classmethod
    def from_okapi(cls, bm25, device: Optional[int] = None) -> "Bm25Index":
        vocab, term_ptr, post_doc, post_tf, doc_len, idf = invert_okapi(bm25)
        return cls(term_ptr, post_doc, post_tf, doc_len, idf, bm25.k1, bm25.b, bm25.avgdl,
                   vocab=vocab, device=device)

    def reweight(self, post_tf, doc_len, idf, k1: float, b: float, avgdl: float) -> None:
        """New k1 / b / avgdl / idf on the same postings (one step of a parameter sweep)."""
        def prep(x, dtype):
            return x.contiguous() if hasattr(x, "data_ptr") else np.ascontiguousarray(x, dtype=dtype)
        post_tf, doc_len, idf = prep(post_tf, np.int32), prep(doc_len, np.int32), prep(idf, np.float64)
        native.call("anr_bm25_reweight", context(self.ctx_device).handle, self.handle,
                    native.ptr(post_tf), native.ptr(doc_len), native.ptr(idf), float(k1), float(b),
                    float(avgdl))

    # -- queries --------------------------------------------------------------
    def term_ids(self, tokens: Iterable[str]) -> np.ndarray:
        """Token strings -> term ids in query order, -1 for tokens absent from the vocabulary."""
        if self.vocab is None:
            raise ValueError("index was built from term ids; pass term ids")
        get = self.vocab.get
        return np.fromiter((get(t, -1) for t in tokens), dtype=np.int32)

    @staticmethod
    def pack_queries(queries: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
        offsets = np.zeros(len(queries) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum([len(q) for q in queries])
        terms = (np.concatenate([np.asarray(q, dtype=np.int32) for q in queries])
                 if offsets[-1] else np.zeros(0, dtype=np.int32))
        return np.ascontiguousarray(terms, dtype=np.int32), offsets

    def search(self, queries: Sequence[Sequence[int]], k: int,
               doc_mask: Optional[np.ndarray] = None, id_base: int = 0):
        """queries: term-id lists.  -> (scores f32 [b,k], docs i32 [b,k], counts i32 [b])."""
        terms, offsets = self.pack_queries(queries)
        b = len(queries)
        scores = np.empty((b, k), dtype=np.float32)
        docs = np.empty((b, k), dtype=np.int32)
        counts = np.empty(b, dtype=np.int32)
        native.call("anr_bm25_search", context(self.ctx_device).handle, self.handle,
                    native.ptr(terms), native.ptr(offsets), b, int(k), native.ptr(doc_mask), None,
                    int(id_base), native.ptr(scores), native.ptr(docs), native.ptr(counts), None)
        return scores, docs, counts

    def search_device(self, queries: Sequence[Sequence[int]], k: int, doc_mask=None, id_base: int = 0):
        """``search`` with the results left on the device: -> (scores f32 [b, k], docs i32 [b, k],
        counts i32 [b]) torch tensors on this index' GPU, enqueued on torch's current stream."""
        import torch
        terms, offsets = self.pack_queries(queries)
        b = len(queries)
        dev = torch.device("cuda", self.ctx_device)
        t_dev = torch.from_numpy(terms if terms.size else np.zeros(1, dtype=np.int32)).to(dev)
        o_dev = torch.from_numpy(offsets).to(dev)
        scores = torch.empty((b, k), dtype=torch.float32, device=dev)
        docs = torch.empty((b, k), dtype=torch.int32, device=dev)
        counts = torch.empty((b,), dtype=torch.int32, device=dev)
        native.call("anr_bm25_search", context(self.ctx_device).handle, self.handle, t_dev.data_ptr(),
                    o_dev.data_ptr(), b, int(k), native.ptr(doc_mask), None, int(id_base),
                    scores.data_ptr(), docs.data_ptr(), counts.data_ptr(), torch_stream_ptr())
        self._keep = (t_dev, o_dev)      # alive until the next call: the kernels read them
        return scores, docs, counts

    def scores(self, term_ids: Sequence[int]) -> np.ndarray:
        """fp32 score of every document for one query (BM25Okapi.get_scores on the device)."""
        terms = np.ascontiguousarray(term_ids, dtype=np.int32)
        out = np.zeros(self.n_docs, dtype=np.float32)
        native.call("anr_bm25_scores", context(self.ctx_device).handle, self.handle,
                    native.ptr(terms), int(terms.shape[0]), native.ptr(out), None)
        return out


# ---------------------------------------------------------------------------------
# fusion
# ---------------------------------------------------------------------------------
WRRF_MAX_ENTRIES = 1 << 22


def wrrf_fuse(id_lists: Sequence[Sequence[int]], weights: Sequence[float], rrf_k: float,
              top_n: Optional[int] = None, device: Optional[int] = None):
    """Weighted RRF of integer-id ranked lists for ONE query -> (ids i32, scores f64)."""
    n_lists = len(id_lists)
    stride = max(1, max((len(l) for l in id_lists), default=1))
    ids = np.full((1, n_lists, stride), -1, dtype=np.int32)
    lens = np.zeros((1, n_lists), dtype=np.int32)
    for i, l in enumerate(id_lists):
        ids[0, i, :len(l)] = l
        lens[0, i] = len(l)
    total = int(lens.sum())
    top_n = max(1, total if top_n is None else min(int(top_n), max(total, 1)))
    w = np.ascontiguousarray(weights, dtype=np.float64)
    out_ids = np.empty((1, top_n), dtype=np.int32)
    out_scores = np.empty((1, top_n), dtype=np.float64)
    out_counts = np.empty(1, dtype=np.int32)
    native.call("anr_wrrf_fuse", context(device).handle, native.ptr(ids), native.ptr(lens),
                native.ptr(w), n_lists, stride, 1, float(rrf_k), top_n, native.ptr(out_ids),
                native.ptr(out_scores), native.ptr(out_counts), None)
    c = int(out_counts[0])
    return out_ids[0, :c], out_scores[0, :c]


def hybrid_search(dense: DenseIndex, bm25: Bm25Index, queries, term_queries, k_dense: int,
                  k_bm25: int, w_dense: float, w_bm25: float, rrf_k: float, top_n: int,
                  row_mask=None, doc_mask=None, doc_to_id=None, id_base: int = 0,
                  want_lists: bool = False):
    """Batched dense + BM25 + WRRF (an extension: the reference is strictly batch-1).

    -> dict(ids i32 [b, top_n], scores f64 [b, top_n], counts i32 [b] [, dense_rows,
    dense_scores, bm25_ids, bm25_scores]).  ``doc_to_id`` must be a device int32 tensor
    (BM25 doc index -> dense row id) or None when both indices share one id space.
    """
    q = _as_f32_matrix(queries)
    b = q.shape[0]
    terms, offsets = Bm25Index.pack_queries(term_queries)
    ids = np.empty((b, top_n), dtype=np.int32)
    scores = np.empty((b, top_n), dtype=np.float64)
    counts = np.empty(b, dtype=np.int32)
    extra = {}
    if want_lists:
        extra = dict(dense_rows=np.empty((b, k_dense), dtype=np.int32),
                     dense_scores=np.empty((b, k_dense), dtype=np.float32),
                     bm25_ids=np.empty((b, k_bm25), dtype=np.int32),
                     bm25_scores=np.empty((b, k_bm25), dtype=np.float32))
    native.call("anr_hybrid_search", context(dense.ctx_device).handle, dense.handle, bm25.handle,
                native.ptr(q), native.ptr(terms), native.ptr(offsets), b, int(k_dense),
                int(k_bm25), native.ptr(row_mask), native.ptr(doc_mask), native.ptr(doc_to_id),
                int(id_base), float(w_dense), float(w_bm25), float(rrf_k), int(top_n),
                native.ptr(ids), native.ptr(scores), native.ptr(counts),
                native.ptr(extra.get("dense_rows")), native.ptr(extra.get("dense_scores")),
                native.ptr(extra.get("bm25_ids")), native.ptr(extra.get("bm25_scores")), None)
    return dict(ids=ids, scores=scores, counts=counts, **extra)
