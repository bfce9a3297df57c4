"""Drop-in ``SearchEngine`` whose retrieval runs on the B200.

Same class name, constructor, method names, positional signatures, defaults, return
types and error convention as the reference's ``src/search_engine.py`` (SURVEY.md
8(b)); the arithmetic -- inner-product scan + top-k (:57-146), BM25 scoring + top-k
(:205-243), weighted RRF (:21-34), source-prefix filtering (:36-55 as a row mask) --
is executed by the CUDA library behind ``include/anr_b200.h``.  There is no CPU
fallback: without the library the import fails, without a B200 the calls raise.
"""
from __future__ import annotations

import functools
import inspect
import logging
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import engine, native, registry

native.load()  # fail at import time, loudly, when the CUDA library has not been built

try:  # the reference imports its tokeniser at module level (search_engine.py:10)
    from processing.preprocess_bm25 import preprocess_text  # reference-style sys.path layout
except Exception:  # pragma: no cover - depends on the caller's sys.path
    from .processing.preprocess_bm25 import preprocess_text


def _fatal(exc: BaseException) -> bool:
    """Errors that must not be folded into the reference's "log and return empty" contract:
    a missing device / library is a deployment fault, not a bad query."""
    return isinstance(exc, ImportError) or (
        isinstance(exc, native.AnrError) and exc.code == 3)


def _never_raises(empty, what: str):
    """The reference's convention for its search methods (search_engine.py:94-98, :144-146,
    :267-269, :291-293): log the error, hand back an empty result.  ``what`` may name the
    ``model_name`` argument of the call.  Deployment faults (``_fatal``) pass through."""
    def wrap(method):
        params = inspect.signature(method)

        @functools.wraps(method)
        def guarded(self, *args, **kwargs):
            try:
                return method(self, *args, **kwargs)
            except Exception as e:
                if _fatal(e):
                    raise
                try:
                    bound = params.bind(self, *args, **kwargs)
                    bound.apply_defaults()
                    label = what.format(**bound.arguments)
                except Exception:
                    label = what
                self.logger.error(f"Error in {label}: {e}")
                return empty()
        return guarded
    return wrap


_NOTHING_LEFT = "No documents found after filtering by filename type: {}"


class SearchEngine:

    def __init__(self, voyage_client, openai_client=None):
        self.vo = voyage_client
        self.openai_client = openai_client
        self.logger = logging.getLogger(__name__)

    # ------------------------------------------------------------------ fusion
    def weighted_reciprocal_rank_fusion(
        self, ranked_lists: List[Tuple], model_weights: Dict[str, float], k: int = 50
    ) -> List[Tuple]:
        """Weighted RRF over several ranked id lists (search_engine.py:21-34) on the device.

        Ids are mapped to dense int32 codes in first-seen order on the host (strings never
        reach the GPU); scores are float64 and bit-identical to the Python loop, order is
        score descending with ties in first-insertion order (Python's stable sort).
        """
        codes: Dict[Hashable, int] = {}
        names: List[Hashable] = []
        id_lists: List[List[int]] = []
        weights: List[float] = []
        for ranked_list, model_name in ranked_lists:
            weights.append(float(model_weights.get(model_name, 1.0)))
            coded = []
            for doc_id in ranked_list:
                c = codes.get(doc_id)
                if c is None:
                    c = codes[doc_id] = len(names)
                    names.append(doc_id)
                coded.append(c)
            id_lists.append(coded)
        if not names:
            return []
        ids, scores = engine.wrrf_fuse(id_lists, weights, float(k))
        return [(names[int(i)], float(s)) for i, s in zip(ids, scores)]

    # ------------------------------------------------------------------ filter
    def _filter_by_filename_type(self, df: pd.DataFrame, filename_type_filter: str) -> pd.DataFrame:
        """Rows whose upper-cased ``source`` starts with any comma-separated prefix (:36-55).

        Kept for callers that use it directly; the searches below do not materialise the
        filtered frame, they pass the same predicate to the kernels as a row bit mask.
        """
        mask = engine.prefix_mask(df["source"].to_numpy(dtype=object), filename_type_filter,
                                  frame_semantics=True)
        filtered_df = df[mask].copy()
        prefix_str = ", ".join(engine.parse_prefixes(filename_type_filter))
        self.logger.info(
            f"Filtered by filename type(s) '{prefix_str}': {len(filtered_df)} documents "
            f"remaining from {len(df)} total"
        )
        return filtered_df

    # ------------------------------------------------------------------ dense
    def _dense_topk(self, query_embedding: np.ndarray, df: pd.DataFrame, similarity_k: int,
                    filename_type_filter: Optional[str]) -> Optional[pd.DataFrame]:
        """Shared body of the two similarity searches.  Returns None when nothing is left
        after filtering (the caller logs and returns the empty filtered frame)."""
        if query_embedding.ndim != 2 or query_embedding.shape[0] != 1:
            # the reference is strictly batch-1: more rows make its flatten()/iloc fail
            raise ValueError(f"expected one query vector, got shape {query_embedding.shape}")
        entry, subset = registry.resolve_frame(df)
        n_rows = len(df)
        mask_words = None
        positions_of = None          # original row -> position in df (None = identity)
        eligible = n_rows
        if subset is not None:
            sub = np.zeros(entry.n, dtype=bool)
            sub[subset] = True
            if filename_type_filter:
                sub &= entry.filter_mask(filename_type_filter)[0]
            eligible = int(sub.sum())
            mask_words = registry.device_words(engine.pack_mask(sub))
            positions_of = pd.Index(subset)
        elif filename_type_filter:
            _, mask_words, eligible = entry.filter_mask(filename_type_filter)
        if eligible == 0:
            return None
        k = max(1, min(int(similarity_k), eligible))
        scores, rows, counts = entry.index().search(query_embedding, k, row_mask=mask_words)
        c = int(counts[0])
        rows, scores = rows[0, :c].astype(np.int64), scores[0, :c]
        pos = rows if positions_of is None else positions_of.get_indexer(rows)
        result_df = df.take(pos)      # = df.iloc[pos].copy() (:89, :137) without the second copy
        similarity = scores.astype(np.result_type(query_embedding.dtype, np.float32))
        if "similarity" in result_df.columns:     # (a frame that already went through a search)
            result_df["similarity"] = similarity
        else:                                     # same frame as the assignment, ~35 us cheaper
            result_df.insert(len(result_df.columns), "similarity", similarity)
        return result_df

    @_never_raises(pd.DataFrame, "{model_name} similarity search with precalculated embedding")
    def similarity_search_with_embedding(self, query_embedding: np.ndarray, df: pd.DataFrame,
                                         model_name: str = "voyage-3-large", similarity_k: int = 25,
                                         filename_type_filter: Optional[str] = None) -> pd.DataFrame:
        """Top-``similarity_k`` rows of ``df`` by inner product with a ready-made query vector, best
        first, plus a ``similarity`` column (search_engine.py:57-98)."""
        if df.empty:
            self.logger.warning(_NOTHING_LEFT.format(filename_type_filter))
            return df
        vector = np.asarray(query_embedding)
        if vector.ndim == 1:
            vector = vector.reshape(1, -1)
        hits = self._dense_topk(vector, df, similarity_k, filename_type_filter)
        if hits is None:
            self.logger.warning(_NOTHING_LEFT.format(filename_type_filter))
            return df.iloc[0:0].copy()
        return hits

    @_never_raises(pd.DataFrame, "{model_name} similarity search")
    def similarity_search(self, query_text: str, df: pd.DataFrame, model_name: str = "voyage-3-large",
                          similarity_k: int = 25, filename_type_filter: Optional[str] = None,
                          query_embedding: Optional[np.ndarray] = None) -> pd.DataFrame:
        """Same search from the query TEXT: the vector is the caller's, if given, else it is
        requested from the embedding service (search_engine.py:100-146).

        A vector that comes back from Voyage is float64 in the reference (:157); it is cast to fp32
        here (documented deviation, <= 1e-7 absolute on unit vectors).
        """
        if df.empty:
            self.logger.warning(_NOTHING_LEFT.format(filename_type_filter))
            return df
        if query_embedding is None:
            vector = self._generate_query_embedding(query_text, model_name)
            self.logger.info(f"Generated new {model_name} query embedding")
        else:
            vector = np.asarray(query_embedding).reshape(1, -1)
            self.logger.info(f"Using provided pre-calculated {model_name} query embedding")
        hits = self._dense_topk(vector, df, similarity_k, filename_type_filter)
        if hits is None:
            self.logger.warning(_NOTHING_LEFT.format(filename_type_filter))
            return df.iloc[0:0].copy()
        self.logger.info(f"{model_name} similarity search found {len(hits)} results")
        return hits

    def _generate_query_embedding(self, query_text: str, model_name: str) -> np.ndarray:
        """Embedding service call (Voyage HTTPS), outside the hot path and unchanged in meaning from
        search_engine.py:148-159: one (1, 2048) row for ``voyage-3-large``, ValueError otherwise."""
        if model_name != "voyage-3-large":
            raise ValueError(f"Unsupported model: {model_name}")
        if not self.vo:
            raise ValueError("Voyage client not available")
        reply = self.vo.embed(query_text, model=model_name, input_type="query", output_dimension=2048)
        return np.array(reply.embeddings).reshape(1, -1)

    def rerank_documents(self, query_text: str, documents: List, reranker_model: str = "rerank-2",
                         reranker_top_k: Optional[int] = None) -> List:
        """Reranker service call (Voyage), outside the hot path; contract of search_engine.py
        :161-203: the documents reordered with a ``rerank_score`` each, or the input unchanged when
        the service fails."""
        if not documents:
            return documents
        try:
            passages = [d.get("document", "") for d in documents]
            self.logger.info(
                f"Starting reranking with model '{reranker_model}' for {len(passages)} documents")
            reply = self.vo.rerank(query=query_text, documents=passages, model=reranker_model,
                                   top_k=reranker_top_k or len(passages), truncation=True)
            ranked = [dict(documents[hit.index], rerank_score=hit.relevance_score)
                      for hit in reply.results if hit.index < len(documents)]
            self.logger.info(f"Reranking completed: {len(ranked)} documents reordered by relevance")
            return ranked
        except Exception as e:
            self.logger.warning(f"Reranking failed, returning original order: {e}")
            return documents

    # ------------------------------------------------------------------ BM25
    def _core_bm25_search(
        self,
        query_tokens: List[str],
        bm25,
        bm25_sections,
        bm25_section_ids,
        similarity_k: int,
        filename_type_filter: Optional[str],
    ) -> List[str]:
        """BM25 top-k section ids, best first (:205-243)."""
        if not query_tokens:
            return []
        entry = registry.resolve_bm25(bm25)
        index = entry.index
        mask_words = None
        eligible = index.n_docs
        if filename_type_filter:
            _, mask_words, eligible = entry.filter_mask(bm25_sections, filename_type_filter)
        if eligible == 0:
            return []
        k = max(1, min(int(similarity_k), eligible))
        terms = index.term_ids(query_tokens)
        _, docs, counts = index.search([terms], k, doc_mask=mask_words)
        return [bm25_section_ids[int(i)] for i in docs[0, :int(counts[0])]]

    @_never_raises(list, "BM25 search")
    def bm25_search(self, query_text: str, bm25, bm25_sections, bm25_section_ids,
                    similarity_k: int = 25, filename_type_filter: Optional[str] = None,
                    use_lemmatized: bool = True) -> List[str]:
        """BM25 top-k section ids for a query TEXT (tokenised like the index was;
        search_engine.py:245-269)."""
        tokens = preprocess_text(query_text, use_lemmatization=use_lemmatized)
        return self._core_bm25_search(tokens, bm25, bm25_sections, bm25_section_ids, similarity_k,
                                      filename_type_filter)

    @_never_raises(list, "preprocessed BM25 search")
    def bm25_search_preprocessed(self, query_tokens: List[str], bm25, bm25_sections, bm25_section_ids,
                                 similarity_k: int = 25,
                                 filename_type_filter: Optional[str] = None) -> List[str]:
        """BM25 top-k section ids for an already tokenised query (search_engine.py:271-293)."""
        return self._core_bm25_search(query_tokens, bm25, bm25_sections, bm25_section_ids,
                                      similarity_k, filename_type_filter)

    # ------------------------------------------------------------------ batched extension
    def hybrid_search_batch(
        self,
        query_embeddings: np.ndarray,
        query_tokens: Sequence[Sequence[str]],
        df: pd.DataFrame,
        bm25,
        bm25_sections,
        bm25_section_ids,
        model_weights: Dict[str, float],
        model_name: str = "voyage-3-large",
        similarity_k: int = 25,
        common_sections_n: int = 25,
        wrrf_k: int = 40,
    ) -> List[List[Tuple[str, float]]]:
        """B queries at once: dense top-k + BM25 top-k + WRRF in ONE library call.

        An extension (the reference is batch-1): equals, per query, the list
        ``weighted_reciprocal_rank_fusion([(dense ids, model_name), (bm25 ids, "BM25")],
        model_weights, wrrf_k)[:common_sections_n]`` of RetrievalEvaluationSystem
        .retrieve_documents (query_rag_retrieval.py:206-212, :308-315, :357-362).
        """
        entry, subset = registry.resolve_frame(df)
        if subset is not None:
            raise ValueError("hybrid_search_batch needs the frame returned by the loader")
        b_entry = registry.resolve_bm25(bm25)
        doc_to_id, extra_ids = _doc_to_row_map(entry, b_entry, bm25_section_ids)
        term_queries = [b_entry.index.term_ids(toks) for toks in query_tokens]
        k_d = max(1, min(int(similarity_k), entry.n))
        k_b = max(1, min(int(similarity_k), b_entry.index.n_docs))
        top_n = max(1, min(int(common_sections_n), k_d + k_b))
        res = engine.hybrid_search(
            entry.index(), b_entry.index, query_embeddings, term_queries, k_d, k_b,
            float(model_weights.get(model_name, 1.0)), float(model_weights.get("BM25", 1.0)),
            float(wrrf_k), top_n, doc_to_id=doc_to_id)
        n_rows = entry.n
        out = []
        for q in range(len(term_queries)):
            c = int(res["counts"][q])
            fused = []
            for i, s in zip(res["ids"][q, :c], res["scores"][q, :c]):
                i = int(i)
                name = entry.ids[i] if i < n_rows else extra_ids[i - n_rows]
                fused.append((name, float(s)))
            out.append(fused)
        return out


def _doc_to_row_map(entry: "registry.DenseEntry", b_entry: "registry.Bm25Entry", section_ids):
    """Device int32 map: BM25 doc index -> id in the dense frame's row space.

    The two stores join on the chunk id string (SURVEY.md section 7 "Common id space"):
    a section whose id is a row of the frame maps to that row, any other section gets a
    fresh id n_rows + j.  Built once per (frame, bm25) pair and cached on the BM25 entry.
    """
    import torch
    cache = b_entry.__dict__.setdefault("_row_maps", {})
    hit = cache.get(entry.key)
    if hit is not None:
        return hit
    row_of = {cid: i for i, cid in enumerate(entry.ids)}
    extra: List[str] = []
    mapping = np.empty(len(section_ids), dtype=np.int32)
    for j, cid in enumerate(section_ids):
        r = row_of.get(cid)
        if r is None:
            r = entry.n + len(extra)
            extra.append(cid)
        mapping[j] = r
    dev = torch.from_numpy(mapping).to(f"cuda:{engine.current_device()}")
    cache[entry.key] = (dev, extra)
    return cache[entry.key]
