"""Batched orchestrator: ``retrieve_documents`` of the reference for B queries at once.

The reference's ``RetrievalEvaluationSystem.retrieve_documents``
(src/query_rag_retrieval.py:149-411) answers ONE query: up to four dense retrievers
(:194-301, in that order), BM25 (:304-343), weighted RRF (:355-367), the top
``common_sections_n`` ids (:369-377), optional reranking (:380-393).  Its evaluator calls it
in a Python loop over thousands of queries (src/retrieval_eval.py:366-378), often with
``similarity_k >= N`` (full ranking, retrieval_eval.py:142).  ``retrieve_documents_batch`` runs
the same pipeline for a batch:

* every dense retriever is ONE ``anr_dense_search`` call over all queries (the tensor-core GEMM
  path for > 32 queries), BM25 ONE ``anr_bm25_search`` call, the fusion ONE ``anr_wrrf_fuse``
  call over ``[B, n_lists, k]`` integer ids (float64, bit-identical to the Python loop); the
  ranked lists STAY ON THE DEVICE between the searches and the fusion (at the evaluator's
  ``similarity_k = 12000`` a batch of 2 048 queries is 100 MB per list: only the fused top-n
  ids, and the hit rows when document dicts are asked for, come back to the host);
* chunk-id strings are mapped to a common int32 id space on the host once per set of frames and
  never reach the GPU;
* per query the result equals what ``retrieve_documents`` returns for that query (a list of
  section ids, or the document dicts with ``return_docs=True``), including its quirks: a
  retriever that returns nothing for a query contributes no list, a single list is passed
  through un-fused, ids are looked up in the documents collected from the lists.

``system`` is duck-typed like the reference class: ``embeddings_data[InfoSource] -> {model: df}``,
``bm25_data[InfoSource] -> (bm25, sections, section_ids)``, ``config.DEFAULT_MODEL_WEIGHTS``,
``search_engine``.  Reranking is a network call in the reference (Voyage); when requested it is
applied per query through ``system.search_engine.rerank_documents`` exactly as :380-385 does.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine, native, registry
from .config import InfoSource

logger = logging.getLogger(__name__)

# order in which retrieve_documents consults the dense retrievers (:194, :220, :250, :277)
DENSE_MODELS = ("voyage-3-large", "voyage-3.5", "text-embedding-3-large", "Qwen3")


def _validate(query_embeddings, similarity_k, common_sections_n, info_source) -> int:
    """query_rag_retrieval.py:94-120, plus: every model carries the same number of queries."""
    if not query_embeddings:
        raise ValueError("Query embeddings dictionary cannot be empty")
    n_queries = None
    for model_name, embedding in query_embeddings.items():
        if not isinstance(embedding, np.ndarray):
            raise ValueError(f"Embedding for {model_name} must be a numpy array")
        if embedding.size == 0:
            raise ValueError(f"Embedding for {model_name} cannot be empty")
        if embedding.ndim != 2:
            raise ValueError(f"Embedding for {model_name} must be a [n_queries, d] matrix")
        if n_queries is None:
            n_queries = embedding.shape[0]
        elif embedding.shape[0] != n_queries:
            raise ValueError("every model must carry one embedding per query")
    if similarity_k <= 0 or common_sections_n <= 0:
        raise ValueError("similarity_k and common_sections_n must be positive integers")
    try:
        InfoSource(info_source.lower())
    except ValueError:
        valid_sources = [s.value for s in InfoSource]
        raise ValueError(f"Invalid info_source '{info_source}'. Must be one of: {valid_sources}")
    return int(n_queries)


class _IdSpace:
    """Common int32 id space of the chunk-id strings of several stores (first seen first)."""

    def __init__(self):
        self.code_of: Dict[str, int] = {}
        self.names: List[str] = []
        self._names_arr: Optional[np.ndarray] = None

    def names_array(self) -> np.ndarray:
        """`names` as an object array (rebuilt when the space has grown): codes -> id strings by
        ONE fancy index per query instead of a Python loop over 12 000 codes (the evaluator's
        full ranking: 3.8 -> 0.34 ms per query on the host)."""
        if self._names_arr is None or len(self._names_arr) != len(self.names):
            arr = np.empty(len(self.names), dtype=object)
            arr[:] = self.names
            self._names_arr = arr
        return self._names_arr

    def codes(self, ids: Sequence[str]) -> np.ndarray:
        out = np.empty(len(ids), dtype=np.int32)
        code_of, names = self.code_of, self.names
        for i, cid in enumerate(ids):
            c = code_of.get(cid)
            if c is None:
                c = code_of[cid] = len(names)
                names.append(cid)
            out[i] = c
        return out


def _store_codes(system, source_enum, key, ids: Sequence[str]) -> Tuple["_IdSpace", np.ndarray]:
    """int32 codes of one store's chunk ids in the id space shared by the stores of a source;
    computed once per store (a pass over N Python strings) and kept on the system object."""
    cache = system.__dict__.setdefault("_anr_id_spaces", {})
    space, codes = cache.setdefault(source_enum, (_IdSpace(), {}))
    hit = codes.get(key)
    if hit is None or len(hit) != len(ids):
        hit = codes[key] = space.codes(ids)
    return space, hit


def _wrrf_batch(ids_dev, lens_dev, weights: Sequence[float], rrf_k: float,
                top_n: int) -> Tuple[np.ndarray, np.ndarray]:
    """ids [B, n_lists, stride] int32, lens [B, n_lists] int32 (device tensors)
    -> (fused ids [B, top_n], counts [B]) on the host."""
    b, n_lists, stride = (int(x) for x in ids_dev.shape)
    out_ids = np.empty((b, top_n), dtype=np.int32)
    out_scores = np.empty((b, top_n), dtype=np.float64)
    out_counts = np.empty(b, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    native.call("anr_wrrf_fuse", engine.context(None).handle, ids_dev.data_ptr(), lens_dev.data_ptr(),
                native.ptr(w), n_lists, stride, b, float(rrf_k), int(top_n), native.ptr(out_ids),
                native.ptr(out_scores), native.ptr(out_counts), engine.torch_stream_ptr())
    return out_ids, out_counts


def _codes_on_device(codes_host: np.ndarray, cache: dict, key):
    """A store's id codes as a resident int32 device tensor (uploaded once per store)."""
    import torch
    hit = cache.get(key)
    if hit is None or hit.numel() != len(codes_host):
        hit = cache[key] = torch.from_numpy(np.ascontiguousarray(codes_host)).to(
            f"cuda:{engine.current_device()}")
    return hit


# queries per device pass: bounds the resident lists to ~0.5 GB per retriever
def _chunk_queries(n_queries: int, k: int) -> int:
    return max(1, min(n_queries, (1 << 27) // max(k, 1)))


def retrieve_documents_batch(
    system,
    query_embeddings: Dict[str, np.ndarray],
    query_texts: Optional[Sequence[str]] = None,
    query_tokens: Optional[Sequence[Sequence[str]]] = None,
    similarity_k: int = 25,
    common_sections_n: int = 15,
    info_source: str = "NICE",
    model_weights: Optional[Dict[str, float]] = None,
    filename_type_filter: Optional[str] = None,
    use_hybrid_search: bool = False,
    wrrf_k: int = 60,
    use_reranker: bool = False,
    reranker_model: str = "rerank-2-lite",
    reranker_top_k: Optional[int] = 5,
    return_docs: bool = False,
) -> List[List]:
    """One entry per query: what ``retrieve_documents`` returns for it (see module docstring)."""
    n_queries = _validate(query_embeddings, similarity_k, common_sections_n, info_source)
    if model_weights is None:
        model_weights = system.config.DEFAULT_MODEL_WEIGHTS.copy()
    source_enum = InfoSource(info_source.lower())
    embeddings_dict = system.embeddings_data.get(source_enum, {})
    bm25_tuple = system.bm25_data.get(source_enum)
    if not embeddings_dict:
        logger.warning(f"No embedding data available for source: {info_source}")
        return [[] for _ in range(n_queries)]
    bm25, bm25_sections, bm25_section_ids = bm25_tuple if bm25_tuple else (None, [], [])

    import torch
    dev = torch.device("cuda", engine.current_device())
    dev_codes = system.__dict__.setdefault("_anr_dev_codes", {})
    want_hits = return_docs or use_reranker

    # ---- which retrievers take part (the same for every query) ---------------------------------
    space = None
    dense_plan = []       # (model, entry, df, mask words, k, frame codes on the device)
    for model in DENSE_MODELS:
        df = embeddings_dict.get(model)
        if df is None or df.empty or model_weights.get(model, 0) <= 0 or model not in query_embeddings:
            continue
        entry, subset = registry.resolve_frame(df)
        if subset is not None:
            raise ValueError("retrieve_documents_batch needs the frames returned by the loader")
        mask_words, eligible = None, entry.n
        if filename_type_filter:
            _, mask_words, eligible = entry.filter_mask(filename_type_filter)
        if eligible == 0:
            continue   # every query: "No documents found after filtering" -> empty result, no list
        k = max(1, min(int(similarity_k), eligible))
        space, frame_codes = _store_codes(system, source_enum, ("dense", entry.key), entry.ids)
        dense_plan.append((model, entry, df, mask_words, k,
                           _codes_on_device(frame_codes, dev_codes, (source_enum, "dense", entry.key))))

    bm25_plan = None      # (index, mask words, k, section codes on the device, token lists)
    if use_hybrid_search and bm25 is not None and model_weights.get("BM25", 0) > 0:
        tokens = None
        if query_tokens is not None:
            tokens = [list(t) if t else [] for t in query_tokens]
        elif query_texts is not None:
            from .search_engine import preprocess_text
            tokens = [preprocess_text(t, use_lemmatization=True) if t else [] for t in query_texts]
        else:
            logger.warning("BM25 search requested but no query_texts or query_tokens provided - skipping BM25")
        if tokens is not None:
            if len(tokens) != n_queries:
                raise ValueError("one token list per query is required")
            b_entry = registry.resolve_bm25(bm25)
            index = b_entry.index
            mask_words, eligible = None, index.n_docs
            if filename_type_filter:
                _, mask_words, eligible = b_entry.filter_mask(bm25_sections, filename_type_filter)
            if eligible > 0:
                k = max(1, min(int(similarity_k), eligible))
                space, sec_codes = _store_codes(system, source_enum, ("bm25", id(b_entry)),
                                                bm25_section_ids)
                bm25_plan = (index, mask_words, k,
                             _codes_on_device(sec_codes, dev_codes, (source_enum, "bm25", id(b_entry))),
                             [index.term_ids(t) for t in tokens])

    n_lists = len(dense_plan) + (1 if bm25_plan else 0)
    if n_lists == 0:
        logger.warning("No ranking methods available - no sections selected")
        return [[] for _ in range(n_queries)]
    list_names = [p[0] for p in dense_plan] + (["BM25"] if bm25_plan else [])
    stride = max([p[4] for p in dense_plan] + ([bm25_plan[2]] if bm25_plan else []))
    weights = [float(model_weights.get(name, 1.0)) for name in list_names]
    top_n = max(1, min(int(common_sections_n), n_lists * stride))

    # ---- searches + weighted RRF, a chunk of queries at a time; lists stay on the device ------
    fused = np.empty((n_queries, top_n), dtype=np.int32)
    fused_counts = np.empty(n_queries, dtype=np.int32)
    dense_hits = [(p[1].ids, p[2], [], [], []) for p in dense_plan]   # + rows, scores, counts (host)
    bm25_host = ([], [])
    step = _chunk_queries(n_queries, stride)
    for q0 in range(0, n_queries, step):
        q1 = min(n_queries, q0 + step)
        b = q1 - q0
        ids_dev = torch.full((b, n_lists, stride), -1, dtype=torch.int32, device=dev)
        lens_dev = torch.zeros((b, n_lists), dtype=torch.int32, device=dev)
        for li, (model, entry, df, mask_words, k, codes_dev) in enumerate(dense_plan):
            q = torch.from_numpy(np.ascontiguousarray(query_embeddings[model][q0:q1],
                                                      dtype=np.float32)).to(dev)
            scores, rows, counts = entry.index().search_device(q, k, row_mask=mask_words)
            ids_dev[:, li, :k] = torch.where(rows >= 0, codes_dev[rows.clamp(min=0).long()],
                                             torch.full_like(rows, -1))
            lens_dev[:, li] = counts
            if want_hits:
                dense_hits[li][2].append(rows.cpu().numpy())
                dense_hits[li][3].append(scores.cpu().numpy())
                dense_hits[li][4].append(counts.cpu().numpy())
        if bm25_plan:
            index, mask_words, k, codes_dev, term_queries = bm25_plan
            # (a query without tokens gets no list: the library returns count 0 for it,
            #  `if not query_tokens: return []`, search_engine.py:216-217)
            _, docs, counts = index.search_device(term_queries[q0:q1], k, doc_mask=mask_words)
            ids_dev[:, n_lists - 1, :k] = torch.where(docs >= 0, codes_dev[docs.clamp(min=0).long()],
                                                      torch.full_like(docs, -1))
            lens_dev[:, n_lists - 1] = counts
            if want_hits:
                bm25_host[0].append(docs.cpu().numpy())
                bm25_host[1].append(counts.cpu().numpy())
        f_ids, f_counts = _wrrf_batch(ids_dev, lens_dev, weights, float(wrrf_k), top_n)
        fused[q0:q1], fused_counts[q0:q1] = f_ids, f_counts
    if want_hits:
        dense_hits = [(ids_, df_, np.concatenate(r), np.concatenate(s), np.concatenate(c))
                      for ids_, df_, r, s, c in dense_hits]
    bm25_docs = ((np.concatenate(bm25_host[0]), np.concatenate(bm25_host[1]))
                 if want_hits and bm25_plan else None)

    names = space.names_array()
    section_of = None
    out: List[List] = []
    for qi in range(n_queries):
        section_ids = names[fused[qi, :int(fused_counts[qi])]].tolist()
        if not return_docs and not use_reranker:
            out.append(section_ids)
            continue
        # the document dicts, first retriever that returned the id wins (:213-218, :240-249, :322-341)
        wanted = set(section_ids)
        docs_by_id: Dict[str, dict] = {}
        for entry_ids, df, rows, scores, counts in dense_hits:
            c = int(counts[qi])
            for r, s in zip(rows[qi, :c], scores[qi, :c]):
                cid = entry_ids[int(r)]
                if cid in wanted and cid not in docs_by_id:
                    rec = df.iloc[int(r)].to_dict()
                    rec["similarity"] = np.float32(s)
                    docs_by_id[cid] = rec
        if bm25_docs is not None:
            if section_of is None:
                section_of = {s.metadata["id"]: s for s in bm25_sections}
            docs, counts = bm25_docs
            for d in docs[qi, :int(counts[qi])]:
                cid = bm25_section_ids[int(d)]
                section = section_of.get(cid)
                if cid in wanted and cid not in docs_by_id and section:
                    docs_by_id[cid] = {"id": cid, "document": section.page_content,
                                       "source": section.metadata.get("source", "Unknown"),
                                       "similarity": 0.0}
        common_docs = [docs_by_id[c] for c in section_ids if c in docs_by_id][:common_sections_n]
        text = query_texts[qi] if query_texts is not None else None
        if use_reranker and common_docs and len(common_docs) > 1 and text:
            common_docs = system.search_engine.rerank_documents(text, common_docs, reranker_model,
                                                                reranker_top_k)
        out.append(common_docs if return_docs else [d.get("id", "Unknown section") for d in common_docs])
    return out
