"""Restatement of ``rank_bm25.BM25Okapi`` (rank-bm25 0.2.2).  TEST INFRASTRUCTURE.

PARITY UNPINNED for the arithmetic: ``rank_bm25`` is a third-party dependency of
the reference (``requirements.txt:5``, unpinned; 0.2.2 is the only modern
release), its source is not under ``/root/reference`` and it cannot be installed
offline.  This file restates the algorithm that release publishes; the anchors
in the reference are the construction call
``BM25Okapi(corpus, k1=1.7, b=0.83, epsilon=0.05)``
(``src/processing/bm25_search.py:77``, params ``:136-139``) and the scoring call
``bm25.get_scores(query_tokens)`` (``src/search_engine.py:219``).
The only published vector for this arithmetic -- the usage example in the package's README,
``get_scores("windy London")`` over three sentences = ``[0., 0.93729472, 0.]`` with the default
parameters -- is reproduced to its 8 printed digits (``tests/test_oracle.py::
test_bm25_okapi_published_known_answer``; on the device in ``tests/test_gpu_parity.py``).  One
vector on a three-document corpus is an anchor, not a pin: the header stays "unpinned".

Attribute names (``corpus_size, avgdl, doc_freqs, idf, doc_len, k1, b, epsilon,
average_idf``) match the package because real pickles written by
``bm25_search.py:84-93`` carry instances of that class and the product's loader
reads exactly those attributes.

Semantics restated (all float64, like the package):
  build : doc_len[i] = len(doc_i); doc_freqs[i] = {term: tf}; nd[term] = number
          of documents containing term; avgdl = sum(doc_len) / corpus_size.
  idf   : idf[t] = ln(N - nd + 0.5) - ln(nd + 0.5); average_idf = mean over the
          vocabulary of the RAW values (negatives included); every term whose
          raw idf is < 0 is then reset to epsilon * average_idf.
  score : for each query token IN ORDER (duplicates repeat):
            score += (idf.get(q) or 0) * tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl))
          with tf = doc_freqs[i].get(q) or 0 per document.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import numpy as np


class BM25Okapi:
    def __init__(self, corpus: Iterable[Sequence[str]], tokenizer=None,
                 k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        self.tokenizer = tokenizer
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        containing = self._count(corpus)
        self._idf_from_counts(containing)

    # -- build --------------------------------------------------------------
    def _count(self, corpus) -> Dict[str, int]:
        containing: Dict[str, int] = {}
        total_tokens = 0
        for doc in corpus:
            self.doc_len.append(len(doc))
            total_tokens += len(doc)
            tf: Dict[str, int] = {}
            for tok in doc:
                tf[tok] = tf.get(tok, 0) + 1
            self.doc_freqs.append(tf)
            for tok in tf:
                containing[tok] = containing.get(tok, 0) + 1
            self.corpus_size += 1
        self.avgdl = total_tokens / self.corpus_size
        return containing

    def _idf_from_counts(self, containing: Dict[str, int]) -> None:
        running = 0
        floored = []
        for tok, nd in containing.items():
            val = math.log(self.corpus_size - nd + 0.5) - math.log(nd + 0.5)
            self.idf[tok] = val
            running += val
            if val < 0:
                floored.append(tok)
        self.average_idf = running / len(self.idf)
        floor = self.epsilon * self.average_idf
        for tok in floored:
            self.idf[tok] = floor

    # -- score --------------------------------------------------------------
    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        out = np.zeros(self.corpus_size)
        dl = np.array(self.doc_len)
        for q in query:
            tf = np.array([(d.get(q) or 0) for d in self.doc_freqs])
            out += (self.idf.get(q) or 0) * (
                tf * (self.k1 + 1) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avgdl))
            )
        return out
