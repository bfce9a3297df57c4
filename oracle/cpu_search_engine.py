"""A CPU ``SearchEngine`` with the reference's DataFrame interface.  TEST INFRASTRUCTURE ONLY.

The reference's own class cannot travel to the GPU box (``/root/reference`` is not there), so the
CPU arm of the evaluator-shaped timing (profiles/f4_evaluator_bench.py) and any test that needs
the reference's STOCK per-query path on DataFrames uses this restatement.  It is assembled from
the functions of ``oracle.retrieval`` (each pinned against the reference module imported verbatim,
tests/test_oracle.py) in the reference's own order of operations:

  similarity_search_with_embedding   src/search_engine.py:57-98   (filter -> np.stack per call ->
                                     np.dot -> argpartition / argsort -> df.iloc[...].copy())
  bm25_search_preprocessed           :271-293 -> _core_bm25_search :205-243 (get_scores, the
                                     filtered branch's Python loop + stable sort)
  weighted_reciprocal_rank_fusion    :21-34

``tests/test_oracle.py::test_cpu_search_engine_equals_the_reference_class`` runs both classes on
the same frames where the reference is mounted.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

from . import retrieval


class CpuSearchEngine:
    def weighted_reciprocal_rank_fusion(self, ranked_lists: List[Tuple],
                                        model_weights: Dict[str, float], k: int = 50) -> List[Tuple]:
        return retrieval.weighted_rrf(ranked_lists, model_weights, k)

    def _filter_by_filename_type(self, df: pd.DataFrame, filename_type_filter: str) -> pd.DataFrame:
        mask = retrieval.filter_mask(df["source"].tolist(), filename_type_filter, frame=True)
        return df[mask].copy()                                            # :48

    def similarity_search_with_embedding(self, query_embedding: np.ndarray, df: pd.DataFrame,
                                         model_name: str = "voyage-3-large", similarity_k: int = 25,
                                         filename_type_filter: Optional[str] = None) -> pd.DataFrame:
        try:
            if filename_type_filter:
                df = self._filter_by_filename_type(df, filename_type_filter)
            if df.empty:
                return df
            embeddings = np.stack(df["embedding"].values)                 # :80, on EVERY call
            similarities = retrieval.dense_scores(query_embedding, embeddings)
            top = retrieval.topk_desc(similarities, similarity_k)
            result_df = df.iloc[top].copy()
            result_df["similarity"] = similarities[top]
            return result_df
        except Exception:
            return pd.DataFrame()

    def bm25_search_preprocessed(self, query_tokens: List[str], bm25, bm25_sections, bm25_section_ids,
                                 similarity_k: int = 25,
                                 filename_type_filter: Optional[str] = None) -> List[str]:
        try:
            if not query_tokens:
                return []
            scores = bm25.get_scores(query_tokens)                        # :219
            sources = ([s.metadata.get("source", "") for s in bm25_sections]
                       if filename_type_filter else None)
            top = retrieval.bm25_topk(scores, similarity_k, sources, filename_type_filter)
            return [bm25_section_ids[int(i)] for i in top]
        except Exception:
            return []
