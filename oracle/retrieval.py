"""numpy restatement of the reference's ``src/search_engine.py`` hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works on plain arrays /
lists instead of DataFrames so it can be applied slab-wise at 1M+ rows.  It is
validated against the reference module imported verbatim in
``tests/test_oracle.py::test_restatement_matches_reference_import`` (runs where
``/root/reference`` exists) and against ``tests/golden/*.npz`` everywhere.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np


# -- search_engine.py:36-55 and :222-231 -------------------------------------
def parse_prefixes(filename_type_filter: str) -> Tuple[str, ...]:
    """``tuple(p.strip().upper() for p in f.split(","))`` (search_engine.py:39, :222)."""
    return tuple(part.strip().upper() for part in filename_type_filter.split(","))


def filter_mask(sources: Sequence[Optional[str]], filename_type_filter: str,
                frame: bool = False) -> np.ndarray:
    """Row mask: upper-cased source starts with any prefix (search_engine.py:40-46).

    ``None``/NaN sources are False (``na=False``).  ``frame=True`` restates the DataFrame
    filter literally: one prefix -> ``startswith`` (:42), several -> the un-escaped regex
    ``^(?:A|B)`` searched in the upper-cased source (:44-46); ``frame=False`` is the BM25
    filter's ``any(startswith)`` (:224-231).  For the alphanumeric prefixes the callers use
    (``"CG,NG"``) the two agree.
    """
    prefixes = parse_prefixes(filename_type_filter)
    out = np.zeros(len(sources), dtype=bool)
    if frame and len(prefixes) > 1:
        import re
        pattern = re.compile("^(?:" + "|".join(prefixes) + ")")
        for i, src in enumerate(sources):
            if isinstance(src, str):
                out[i] = pattern.search(src.upper()) is not None
        return out
    for i, src in enumerate(sources):
        if isinstance(src, str):
            out[i] = src.upper().startswith(prefixes)
    return out


# -- search_engine.py:77-90 / :128-138 ----------------------------------------
def dense_scores(query: np.ndarray, emb: np.ndarray) -> np.ndarray:
    """``np.dot(q.reshape(1,-1), E.T).flatten()`` (search_engine.py:77-81)."""
    q = query.reshape(1, -1) if query.ndim == 1 else query
    return np.dot(q, emb.T).flatten()


def topk_desc(scores: np.ndarray, k: int) -> np.ndarray:
    """Top-k positions, best first (search_engine.py:83-87, :131-135, :237-241)."""
    if len(scores) > k:
        part = np.argpartition(scores, -k)[-k:]
        return part[scores[part].argsort()[::-1]]
    return scores.argsort()[::-1]


def dense_topk(query: np.ndarray, emb: np.ndarray, k: int,
               row_mask: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(row indices into ``emb``, scores), best first, optionally over masked rows only."""
    if row_mask is not None:
        rows = np.flatnonzero(row_mask)
        if rows.size == 0:
            return rows.astype(np.int64), np.zeros(0, dtype=emb.dtype)
        s = dense_scores(query, emb[rows])
        top = topk_desc(s, k)
        return rows[top], s[top]
    s = dense_scores(query, emb)
    top = topk_desc(s, k)
    return top, s[top]


# -- search_engine.py:205-243 -------------------------------------------------
def bm25_topk(scores: np.ndarray, k: int,
              doc_sources: Optional[Sequence[str]] = None,
              filename_type_filter: Optional[str] = None) -> np.ndarray:
    """Doc indices, best first.  Filtered branch = stable descending sort (:224-233)."""
    if filename_type_filter:
        prefixes = parse_prefixes(filename_type_filter)
        kept = [
            (i, scores[i])
            for i, src in enumerate(doc_sources)
            if any((src or "").upper().startswith(p) for p in prefixes)
        ]
        best = sorted(kept, key=lambda t: t[1], reverse=True)[:k]
        return np.array([i for i, _ in best], dtype=np.int64)
    return topk_desc(np.array(scores), k)


# -- search_engine.py:21-34 ---------------------------------------------------
def weighted_rrf(ranked_lists: Sequence[Tuple[Sequence[Hashable], str]],
                 model_weights: Dict[str, float], k: float = 50) -> List[Tuple[Hashable, float]]:
    """score[id] += w * (1 / (k + rank)), rank from 1; stable sort by score desc."""
    acc: Dict[Hashable, float] = {}
    for ids, name in ranked_lists:
        w = model_weights.get(name, 1.0)
        for rank, doc_id in enumerate(ids, start=1):
            acc[doc_id] = acc.get(doc_id, 0.0) + w * (1 / (k + rank))
    return sorted(acc.items(), key=lambda t: t[1], reverse=True)


# -- comparison helpers shared by the parity tests ----------------------------
def assert_ranking_matches(got_ids, got_scores, want_ids, want_scores, all_scores=None,
                           rtol: float = 1e-5, atol: float = 1e-7, what: str = "") -> None:
    """Top-k parity under the tie rule of SURVEY.md 8(c).

    * same length; scores position-wise within ``rtol`` relative (+``atol``);
    * ids identical except inside groups of (near-)equal oracle score: at every
      position the id we got must have an oracle score within tolerance of the
      oracle's score at that position (``all_scores`` = the full oracle score
      vector, so an id outside the oracle's top-k that ties the k-th is legal).
    """
    got_ids = np.asarray(got_ids)
    want_ids = np.asarray(want_ids)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    want_scores = np.asarray(want_scores, dtype=np.float64)
    assert got_ids.shape == want_ids.shape, f"{what}: length {got_ids.shape} != {want_ids.shape}"
    tol = rtol * np.abs(want_scores) + atol
    bad = np.abs(got_scores - want_scores) > tol
    assert not bad.any(), (
        f"{what}: scores differ at {np.flatnonzero(bad)[:5]}: "
        f"{got_scores[bad][:5]} vs {want_scores[bad][:5]}")
    assert len(set(got_ids.tolist())) == len(got_ids), f"{what}: duplicate ids"
    diff = np.flatnonzero(got_ids != want_ids)
    if diff.size == 0:
        return
    assert all_scores is not None, f"{what}: ids differ at {diff[:5]} and no score vector given"
    all_scores = np.asarray(all_scores, dtype=np.float64)
    for p in diff:
        s_have = all_scores[int(got_ids[p])]
        assert abs(s_have - want_scores[p]) <= tol[p], (
            f"{what}: position {p}: id {got_ids[p]} (oracle score {s_have}) is not a tie of "
            f"oracle id {want_ids[p]} (score {want_scores[p]})")
