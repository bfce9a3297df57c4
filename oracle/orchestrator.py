"""CPU restatement of the reference orchestrator for ONE query.  TEST INFRASTRUCTURE ONLY.

Follows ``RetrievalEvaluationSystem.retrieve_documents`` (src/query_rag_retrieval.py:149-411):
dense retrievers in the order voyage-3-large, voyage-3.5, text-embedding-3-large, Qwen3
(:194-301), BM25 (:304-343), weighted RRF when more than one ranked list exists (:355-367),
document lookup and truncation (:369-377), optional reranking (:380-393), the list of section ids
or the document dicts (:399-407); any exception -> ``[]`` (:409-411).

It calls the search methods of ``system.search_engine`` -- the reference's own SearchEngine when
it is pinned against the unmodified reference method (tests/test_oracle.py, needs
/root/reference), the drop-in SearchEngine when it checks ``retrieve_documents_batch`` on the GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

DENSE_MODELS = ("voyage-3-large", "voyage-3.5", "text-embedding-3-large", "Qwen3")


def retrieve_documents(system, info_source_enum, query_embeddings: Dict[str, np.ndarray],
                       query_text: Optional[str] = None, query_tokens: Optional[List[str]] = None,
                       similarity_k: int = 25, common_sections_n: int = 15,
                       model_weights: Optional[Dict[str, float]] = None,
                       filename_type_filter: Optional[str] = None, use_hybrid_search: bool = False,
                       wrrf_k: int = 60, use_reranker: bool = False,
                       reranker_model: str = "rerank-2-lite", reranker_top_k: Optional[int] = 5,
                       return_docs: bool = False):
    if model_weights is None:
        model_weights = system.config.DEFAULT_MODEL_WEIGHTS.copy()
    try:
        frames = system.embeddings_data.get(info_source_enum, {})
        bm25_tuple = system.bm25_data.get(info_source_enum)
        if not frames:
            return []
        bm25, sections, section_ids = bm25_tuple if bm25_tuple else (None, [], [])
        section_of = {s.metadata["id"]: s for s in sections}
        se = system.search_engine
        ranked_lists, collected = [], []          # [(ids, model)], [(id, document dict)]
        for model in DENSE_MODELS:               # :194-301, one block per model in the reference
            df = frames.get(model)
            if df is None or df.empty or model_weights.get(model, 0) <= 0 \
                    or model not in query_embeddings:
                continue
            hits = se.similarity_search_with_embedding(query_embeddings[model], df, model,
                                                       similarity_k, filename_type_filter)
            if hits.empty:
                continue
            ranked_lists.append((hits["id"].tolist(), model))
            seen = {doc_id for doc_id, _ in collected}
            fresh = hits[~hits["id"].isin(seen)]
            collected.extend(zip(fresh["id"], fresh.to_dict("records")))
        if use_hybrid_search and bm25 is not None and model_weights.get("BM25", 0) > 0:   # :304-343
            bm25_ids = None
            if query_tokens:
                bm25_ids = se.bm25_search_preprocessed(query_tokens, bm25, sections, section_ids,
                                                       similarity_k, filename_type_filter)
            elif query_text:
                bm25_ids = se.bm25_search(query_text, bm25, sections, section_ids, similarity_k,
                                          filename_type_filter)
            if bm25_ids:
                ranked_lists.append((bm25_ids, "BM25"))
                seen = {doc_id for doc_id, _ in collected}
                for sid in bm25_ids:
                    section = section_of.get(sid)
                    if sid not in seen and section:
                        collected.append((sid, {"id": sid, "document": section.page_content,
                                                "source": section.metadata.get("source", "Unknown"),
                                                "similarity": 0.0}))
        if len(ranked_lists) > 1:                                                          # :355-367
            fused = se.weighted_reciprocal_rank_fusion(ranked_lists, model_weights, wrrf_k)
            chosen = [sid for sid, _ in fused[:common_sections_n]]
        elif len(ranked_lists) == 1:
            chosen = ranked_lists[0][0][:common_sections_n]
        else:
            chosen = []
        by_id = {doc_id: doc for doc_id, doc in collected}
        docs = [by_id[sid] for sid in chosen if sid in by_id][:common_sections_n] if collected else []
        if use_reranker and docs and len(docs) > 1 and query_text:                         # :380-385
            docs = se.rerank_documents(query_text, docs, reranker_model, reranker_top_k)
        if return_docs:
            return docs
        return [d.get("id", "Unknown section") for d in docs]
    except Exception:
        return []
