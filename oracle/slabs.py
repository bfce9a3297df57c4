"""The oracle applied SLAB-WISE to corpora that do not fit the host (10M+ rows).  TEST
INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by ``bench.py``'s parity legs and by
``tests/`` -- never by the product.

SURVEY.md 8(d) "Parity check in the same run": for N >= 10M the oracle is the same numpy
statements -- ``np.dot`` + ``argpartition`` / ``argsort`` of src/search_engine.py:81-87 -- applied
to slabs of rows streamed from the device, with an exact merge of the per-slab top-k; BM25 is
scored from the postings of the query's own terms only (``oracle.csr.scores`` on a sub-index whose
vocabulary is those terms), and a corpus sharded by chunk is scored shard by shard with the GLOBAL
idf / avgdl (a document's BM25 score depends on its own postings and on corpus-wide statistics
only), the per-shard score vectors being concatenated in document order.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import csr, retrieval


def dense_topk_slabs(slabs: Iterable[Tuple[int, np.ndarray]], queries: np.ndarray, k: int):
    """slabs: (first row, [rows, d] float32) pairs covering the corpus once.
    -> (ids int64 [nq, <=k], scores [nq, <=k]) best first.  Per slab the reference's statements
    (np.dot of every query, top-k by argpartition + argsort); the slab winners are merged by one
    more descending sort (exact: the global top-k is a subset of the union of slab top-ks)."""
    queries = np.asarray(queries, dtype=np.float32).reshape(-1, queries.shape[-1])
    nq = queries.shape[0]
    best_ids: List[List[np.ndarray]] = [[] for _ in range(nq)]
    best_sc: List[List[np.ndarray]] = [[] for _ in range(nq)]
    for row0, emb in slabs:
        for q in range(nq):
            s = retrieval.dense_scores(queries[q], emb)
            top = retrieval.topk_desc(s, k)
            best_ids[q].append(top.astype(np.int64) + int(row0))
            best_sc[q].append(s[top])
    out_ids, out_sc = [], []
    for q in range(nq):
        ids = np.concatenate(best_ids[q]) if best_ids[q] else np.zeros(0, dtype=np.int64)
        sc = np.concatenate(best_sc[q]) if best_sc[q] else np.zeros(0, dtype=np.float32)
        order = retrieval.topk_desc(sc, k)
        out_ids.append(ids[order])
        out_sc.append(sc[order])
    return out_ids, out_sc


def bm25_subindex(fetch_term: Callable[[int], Tuple[np.ndarray, np.ndarray]],
                  term_ids: Sequence[int], doc_len: np.ndarray, idf_of: Callable[[int], float],
                  avgdl: float, k1: float, b: float) -> Tuple[csr.CsrIndex, Dict[int, int]]:
    """A ``CsrIndex`` over the DISTINCT valid terms of ``term_ids`` only (what one query needs of
    a 10M-document index): fetch_term(t) -> (post_doc int32, post_tf int32).  Returns the index
    and the map original term id -> term id inside it."""
    distinct = [t for t in dict.fromkeys(int(t) for t in term_ids) if t >= 0]
    docs, tfs, ptr = [], [], [0]
    for t in distinct:
        d, f = fetch_term(t)
        docs.append(np.asarray(d, dtype=np.int32))
        tfs.append(np.asarray(f, dtype=np.int32))
        ptr.append(ptr[-1] + len(d))
    index = csr.CsrIndex(
        term_ptr=np.asarray(ptr, dtype=np.int64),
        post_doc=np.concatenate(docs) if docs else np.zeros(0, dtype=np.int32),
        post_tf=np.concatenate(tfs) if tfs else np.zeros(0, dtype=np.int32),
        doc_len=np.asarray(doc_len, dtype=np.int64),
        idf=np.array([idf_of(t) for t in distinct], dtype=np.float64),
        avgdl=float(avgdl), k1=float(k1), b=float(b))
    return index, {t: i for i, t in enumerate(distinct)}


def bm25_scores_subindex(fetch_term, term_ids, doc_len, idf_of, avgdl, k1, b) -> np.ndarray:
    """float64 BM25 score of every document for ONE query (duplicates repeat, -1 = unknown)."""
    index, local = bm25_subindex(fetch_term, term_ids, doc_len, idf_of, avgdl, k1, b)
    return csr.scores(index, [local.get(int(t), -1) for t in term_ids])


def assert_topk_matches(got_ids, got_scores, want_ids, want_scores, score_of: Callable[[int], float],
                        rtol: float = 1e-5, atol: float = 1e-6, what: str = "") -> None:
    """``retrieval.assert_ranking_matches`` without the full score vector: ``score_of(id)`` gives
    the oracle's score of one id (needed only where the ids differ, i.e. inside tie groups)."""
    got_ids, want_ids = np.asarray(got_ids), np.asarray(want_ids)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    want_scores = np.asarray(want_scores, dtype=np.float64)
    assert got_ids.shape == want_ids.shape, f"{what}: length {got_ids.shape} != {want_ids.shape}"
    tol = rtol * np.abs(want_scores) + atol
    bad = np.abs(got_scores - want_scores) > tol
    assert not bad.any(), (f"{what}: scores differ at {np.flatnonzero(bad)[:5]}: "
                           f"{got_scores[bad][:5]} vs {want_scores[bad][:5]}")
    assert len(set(got_ids.tolist())) == len(got_ids), f"{what}: duplicate ids"
    for p in np.flatnonzero(got_ids != want_ids):
        s_have = float(score_of(int(got_ids[p])))
        assert abs(s_have - want_scores[p]) <= tol[p], (
            f"{what}: position {p}: id {got_ids[p]} (oracle score {s_have}) is not a tie of "
            f"oracle id {want_ids[p]} (score {want_scores[p]})")
