"""Shared by the CPU and GPU tests: corpora rebuilt from the golden inputs, tie-aware checks."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bm25_okapi, csr, retrieval  # noqa: E402

synth = importlib.import_module("a-nice-rag_b200.synth")

WEIGHTS = {"voyage-3-large": 5.0, "BM25": 1.0}
WRRF_K = 40
K1, B, EPS = 1.7, 0.83, 0.05          # src/processing/bm25_search.py:136-139
RTOL = 1e-5                            # north_star: scores within 1e-5 relative


def tag(flt):
    return "none" if flt is None else flt.replace(",", "_").replace(" ", "")


def okapi_from_case(case):
    corpus = synth.doc_token_lists(case["doc_ptr"], case["tokens"])
    return bm25_okapi.BM25Okapi(corpus, k1=K1, b=B, epsilon=EPS)


def csr_from_case(case):
    """CSR index whose term ids are the Zipf ranks (token "t{r}" -> r); idf taken from the
    literal BM25Okapi build so both agree bit for bit."""
    okapi = okapi_from_case(case)
    vocab = int(case["vocab"])
    ix = csr.from_token_ids(case["doc_ptr"], case["tokens"], vocab, K1, B, EPS)
    idf = np.zeros(vocab, dtype=np.float64)
    for tok, val in okapi.idf.items():
        idf[int(tok[1:])] = val
    ix.idf = idf
    ix.avgdl = float(okapi.avgdl)
    return ix, okapi


def filter_mask(case, flt):
    return retrieval.filter_mask(list(case["sources"]), flt)


def check_topk(got_ids, got_scores, want_ids, want_scores, all_scores, what, rtol=RTOL,
               atol=1e-6):
    retrieval.assert_ranking_matches(got_ids, got_scores, want_ids, want_scores,
                                     all_scores=all_scores, rtol=rtol, atol=atol, what=what)


def check_ids_only(got_ids, want_ids, all_scores, what, rtol=RTOL, atol=1e-6):
    """Ranking parity when only ids are returned (BM25 returns ids): same length, no
    duplicates, and every position holds an id whose oracle score ties the expected one."""
    got_ids = np.asarray(got_ids)
    want_ids = np.asarray(want_ids)
    assert got_ids.shape == want_ids.shape, f"{what}: {got_ids.shape} vs {want_ids.shape}"
    assert len(set(got_ids.tolist())) == len(got_ids), f"{what}: duplicate ids"
    s_got = np.asarray(all_scores, dtype=np.float64)[got_ids]
    s_want = np.asarray(all_scores, dtype=np.float64)[want_ids]
    bad = np.abs(s_got - s_want) > rtol * np.abs(s_want) + atol
    assert not bad.any(), f"{what}: positions {np.flatnonzero(bad)[:5]} are not ties"
