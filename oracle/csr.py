"""Vectorised numpy CSR form of the BM25Okapi scoring.  TEST INFRASTRUCTURE ONLY.

The literal restatement (``oracle/bm25_okapi.py``) does T*N Python dict lookups
per query -- minutes per query at 10M documents.  This module scores the same
formula over an inverted index so the oracle stays usable at BASELINE sizes.
``tests/test_oracle.py`` proves it bit-equal (float64) to the literal form on
small corpora: the per-element expression is evaluated in the same order, and
documents without the term receive ``+ idf * 0.0`` there, i.e. are unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass
class CsrIndex:
    term_ptr: np.ndarray     # int64 [V+1]
    post_doc: np.ndarray     # int32 [nnz], ascending doc id inside each term
    post_tf: np.ndarray      # int32 [nnz]
    doc_len: np.ndarray      # int64 [N]
    idf: np.ndarray          # float64 [V]  (epsilon floor already applied)
    avgdl: float
    k1: float
    b: float
    vocab: Optional[Dict[str, int]] = None   # token string -> term id

    @property
    def n_docs(self) -> int:
        return int(self.doc_len.shape[0])


def idf_table(n_docs: int, nd: np.ndarray, epsilon: float) -> np.ndarray:
    """idf per term in VOCABULARY ORDER, rank-bm25 0.2.2 ``_calc_idf`` semantics.

    The running sum is accumulated in vocabulary order with Python floats, as the
    package does, so ``average_idf`` (and the epsilon floor) is bit-identical.
    """
    raw = [math.log(n_docs - int(c) + 0.5) - math.log(int(c) + 0.5) for c in nd]
    total = 0
    for v in raw:
        total += v
    avg = total / len(raw)
    floor = epsilon * avg
    return np.array([floor if v < 0 else v for v in raw], dtype=np.float64)


def from_okapi(bm25) -> CsrIndex:
    """Invert a BM25Okapi-shaped object (attrs listed in bm25_okapi.py)."""
    vocab = {tok: i for i, tok in enumerate(bm25.idf.keys())}
    lists: List[List[int]] = [[] for _ in vocab]
    tfs: List[List[int]] = [[] for _ in vocab]
    for d, freqs in enumerate(bm25.doc_freqs):
        for tok, tf in freqs.items():
            t = vocab[tok]
            lists[t].append(d)
            tfs[t].append(tf)
    ptr = np.zeros(len(vocab) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(x) for x in lists])
    return CsrIndex(
        term_ptr=ptr,
        post_doc=np.fromiter((d for l in lists for d in l), dtype=np.int32, count=int(ptr[-1])),
        post_tf=np.fromiter((f for l in tfs for f in l), dtype=np.int32, count=int(ptr[-1])),
        doc_len=np.asarray(bm25.doc_len, dtype=np.int64),
        idf=np.array([bm25.idf[t] for t in vocab], dtype=np.float64),
        avgdl=float(bm25.avgdl), k1=float(bm25.k1), b=float(bm25.b), vocab=vocab,
    )


def from_token_ids(doc_ptr: np.ndarray, tokens: np.ndarray, n_vocab: int,
                   k1: float, b: float, epsilon: float) -> CsrIndex:
    """Invert a flat token-id corpus (doc i = tokens[doc_ptr[i]:doc_ptr[i+1]]).

    Term ids with no occurrence get an empty posting list and idf 0 (they are
    "unknown terms": ``idf.get(q) or 0``).  The idf running sum is taken over the
    PRESENT terms in ascending term-id order, which is the vocabulary order a
    ``BM25Okapi`` would have if tokens were first seen in that order; tests that
    compare with the literal class build their corpora accordingly or compare
    idf with a tolerance.
    """
    n_docs = len(doc_ptr) - 1
    doc_len = np.diff(doc_ptr).astype(np.int64)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), doc_len)
    key = tokens.astype(np.int64) * n_docs + doc_of
    uniq, tf = np.unique(key, return_counts=True)
    term = uniq // n_docs
    nd = np.bincount(term, minlength=n_vocab)
    ptr = np.zeros(n_vocab + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(nd)
    present = np.flatnonzero(nd)
    idf = np.zeros(n_vocab, dtype=np.float64)
    idf[present] = idf_table(n_docs, nd[present], epsilon)
    return CsrIndex(
        term_ptr=ptr, post_doc=(uniq % n_docs).astype(np.int32), post_tf=tf.astype(np.int32),
        doc_len=doc_len, idf=idf, avgdl=float(doc_len.sum() / n_docs), k1=k1, b=b,
    )


def scores(index: CsrIndex, term_ids: Sequence[int]) -> np.ndarray:
    """float64 BM25 scores for one query given as term ids (-1 = unknown term)."""
    out = np.zeros(index.n_docs)
    for t in term_ids:
        if t < 0:
            continue
        lo, hi = int(index.term_ptr[t]), int(index.term_ptr[t + 1])
        docs = index.post_doc[lo:hi]
        tf = index.post_tf[lo:hi].astype(np.int64)
        dl = index.doc_len[docs]
        out[docs] += (index.idf[t] or 0) * (
            tf * (index.k1 + 1) / (tf + index.k1 * (1 - index.b + index.b * dl / index.avgdl))
        )
    return out


def scores_for_tokens(index: CsrIndex, tokens: Sequence[str]) -> np.ndarray:
    return scores(index, [index.vocab.get(t, -1) for t in tokens])
