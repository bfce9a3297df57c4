"""End-to-end CPU statement of one hybrid query, on plain arrays.  TEST INFRASTRUCTURE ONLY.

Mirrors the body of ``RetrievalEvaluationSystem.retrieve_documents`` for one dense model +
BM25 (``src/query_rag_retrieval.py:206-212, :308-315, :357-362``): dense top-k ids, BM25
top-k ids, weighted RRF over the two lists, first ``top_n`` fused ids.  Built only from the
functions of ``oracle.retrieval`` / ``oracle.csr`` (each of which cites the reference lines
it restates).  Also the timed CPU baseline of ``bench.py`` (kind "port").
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import csr, retrieval


def hybrid_query(query: np.ndarray, emb: np.ndarray, index: csr.CsrIndex,
                 term_ids: Sequence[int], k_dense: int, k_bm25: int, weights: Dict[str, float],
                 rrf_k: float, top_n: int, doc_to_id: Optional[np.ndarray] = None,
                 model_name: str = "voyage-3-large"):
    """-> dict(fused=[(id, score)], dense_ids, dense_scores, bm25_ids, bm25_scores (float64,
    all docs)).  Ids are dense row numbers; BM25 doc i maps to ``doc_to_id[i]`` (identity if None)."""
    d_rows, d_scores = retrieval.dense_topk(query, emb, k_dense)
    b_all = csr.scores(index, term_ids)
    b_docs = retrieval.bm25_topk(b_all, k_bm25)
    b_ids = b_docs if doc_to_id is None else doc_to_id[b_docs]
    fused = retrieval.weighted_rrf(
        [([int(i) for i in d_rows], model_name), ([int(i) for i in b_ids], "BM25")],
        weights, rrf_k)[:top_n]
    return dict(fused=fused, dense_ids=d_rows, dense_scores=d_scores, bm25_docs=b_docs,
                bm25_ids=b_ids, bm25_all=b_all)


def hybrid_query_sharded(query: np.ndarray, shards: Sequence[Tuple[int, np.ndarray, csr.CsrIndex]],
                         term_ids: Sequence[int], k_dense: int, k_bm25: int,
                         weights: Dict[str, float], rrf_k: float, top_n: int,
                         model_name: str = "voyage-3-large"):
    """The same query over a corpus held as chunk shards ``(first row, emb, index)`` in row order:
    the score vectors of the shards are concatenated (a dense score is per row; a BM25 score is
    per document GIVEN corpus-wide idf / avgdl, which every shard's ``index`` must carry) and the
    reference's top-k + fusion statements run on the concatenation -- what an unsharded reference
    returns on the whole corpus."""
    d_all = np.concatenate([retrieval.dense_scores(query, emb) for _, emb, _ in shards])
    b_all = np.concatenate([csr.scores(ix, term_ids) for _, _, ix in shards])
    d_rows = retrieval.topk_desc(d_all, k_dense)
    b_docs = retrieval.bm25_topk(b_all, k_bm25)
    fused = retrieval.weighted_rrf(
        [([int(i) for i in d_rows], model_name), ([int(i) for i in b_docs], "BM25")],
        weights, rrf_k)[:top_n]
    return dict(fused=fused, dense_ids=d_rows, dense_scores=d_all[d_rows], bm25_docs=b_docs,
                bm25_ids=b_docs, bm25_all=b_all, dense_all=d_all)


def check_fused(got_ids: Sequence[int], got_scores: Sequence[float],
                dense_ids: Sequence[int], bm25_ids: Sequence[int], weights: Tuple[float, float],
                rrf_k: float, top_n: int) -> None:
    """Exact check of a fused list GIVEN the two ranked lists it was built from: float64
    scores bit-identical, order = stable sort (first-insertion order among ties)."""
    want = retrieval.weighted_rrf([(list(dense_ids), "d"), (list(bm25_ids), "b")],
                                  {"d": weights[0], "b": weights[1]}, rrf_k)[:top_n]
    assert [int(i) for i in got_ids] == [int(i) for i, _ in want], (list(got_ids), want)
    assert [float(s) for s in got_scores] == [s for _, s in want], (list(got_scores), want)
