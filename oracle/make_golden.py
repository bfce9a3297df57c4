"""Generates tests/golden/*.npz by running the REFERENCE's own code in this container.

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden`` where ``/root/reference``
exists.  The reference's ``SearchEngine`` / ``DatabaseManager`` are imported verbatim
(``oracle.reference_loader``) and driven through the real on-disk formats (SQLite ``chunks``
table, BM25 pickle) written by the synthetic-data writers; the only stand-in is
``rank_bm25.BM25Okapi`` (absent offline) = ``oracle.bm25_okapi.BM25Okapi``.

What the vectors pin: the reference's dense top-k ids + similarities, BM25 top-k ids (both
branches: argpartition and the filtered stable sort), weighted RRF lists, the k >= N full
ranking, filters, and the loader's DataFrame contract.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

from . import bm25_okapi, reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
synth = importlib.import_module("a-nice-rag_b200.synth")

WEIGHTS = {"voyage-3-large": 5.0, "BM25": 1.0}     # src/config.py:31,35
WRRF_K = 40                                        # src/retrieval_eval.py:279


def small_case_inputs():
    """N=2048 x D=52 (not a multiple of 4), planted exact ties, tiny vocabulary so that head
    terms have negative raw idf; returned as plain arrays (stored in the npz)."""
    n, d, vocab = 2048, 52, 400
    emb = synth.unit_vectors(n, d, seed=11)
    emb[100] = emb[7]; emb[1999] = emb[7]; emb[512] = emb[513]          # exact dense ties
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=12, len_lo=20, len_hi=60)
    # exact BM25 ties: documents 40, 41, 1500 are copies of document 39
    lo, hi = doc_ptr[39], doc_ptr[40]
    docs = [tokens[doc_ptr[i]:doc_ptr[i + 1]].copy() for i in range(n)]
    for j in (40, 41, 1500):
        docs[j] = tokens[lo:hi].copy()
    doc_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(x) for x in docs], out=doc_ptr[1:])
    tokens = np.concatenate(docs).astype(np.int32)
    queries = synth.unit_vectors(12, d, seed=13)
    queries[3] = emb[7]                                                  # query hits the tie group
    tq = synth.zipf_queries(12, 8, vocab, 1.1, seed=14)
    tq[2, 5] = tq[2, 1]                                                  # duplicate term
    tq[4, :] = vocab + 50                                                # unknown terms only
    tq[5, 0] = vocab + 7                                                 # one unknown term
    tq[6, :] = docs[39][:8]                                              # hits the BM25 tie group
    tq[7, :] = 0                                                         # most frequent term x8
    srcs = synth.sources(n, seed=15)
    return dict(emb=emb, doc_ptr=doc_ptr, tokens=tokens, queries=queries, term_queries=tq,
                sources=np.array(srcs, dtype=object), vocab=np.int64(vocab))


def run_reference(case: dict, ks, filters, tmpdir: str) -> dict:
    ref = reference_loader.load_reference()
    n = case["emb"].shape[0]
    srcs = list(case["sources"])
    ids = synth.chunk_ids(n, srcs)
    contents = [f"content {i}" for i in range(n)]
    corpus = synth.doc_token_lists(case["doc_ptr"], case["tokens"])
    bm25 = bm25_okapi.BM25Okapi(corpus, k1=1.7, b=0.83, epsilon=0.05)   # bm25_search.py:136-139
    db_path = os.path.join(tmpdir, "chunks.db")
    pkl_path = os.path.join(tmpdir, "bm25.pkl")
    synth.write_chunks_db(db_path, ids, contents, srcs, case["emb"])
    synth.write_bm25_pickle(pkl_path, bm25, contents, ids, srcs)

    # the reference's loader unpickles rank_bm25.BM25Okapi / langchain Document by name
    rank_mod = types.ModuleType("rank_bm25"); rank_mod.BM25Okapi = bm25_okapi.BM25Okapi
    lc = types.ModuleType("langchain"); lcs = types.ModuleType("langchain.schema")
    lcd = types.ModuleType("langchain.schema.document"); lcd.Document = synth.Document
    injected = {"rank_bm25": rank_mod, "langchain": lc, "langchain.schema": lcs,
                "langchain.schema.document": lcd}
    saved_cls = (bm25_okapi.BM25Okapi.__module__, bm25_okapi.BM25Okapi.__qualname__)
    sys.modules.update(injected)
    bm25_okapi.BM25Okapi.__module__ = "rank_bm25"
    try:
        dm = ref.DatabaseManager()
        df = dm.load_embeddings_from_sql(db_path, "voyage-3-large")
        r_bm25, r_sections, r_section_ids = dm.load_bm25_from_pickle(pkl_path)
    finally:
        bm25_okapi.BM25Okapi.__module__ = saved_cls[0]
        for name in injected:
            sys.modules.pop(name, None)
    assert list(df.columns) == ["id", "document", "source", "embedding", "url"]
    assert list(df["id"]) == ids and r_section_ids == ids
    row_of = {cid: i for i, cid in enumerate(ids)}

    se = ref.SearchEngine(None, None)
    out = {}
    nq = case["queries"].shape[0]
    for flt in filters:
        tag = "none" if flt is None else flt.replace(",", "_").replace(" ", "")
        for k in ks:
            d_ids = np.full((nq, min(k, n)), -1, dtype=np.int64)
            d_sc = np.zeros((nq, min(k, n)), dtype=np.float32)
            d_cnt = np.zeros(nq, dtype=np.int64)
            b_ids = np.full((nq, min(k, n)), -1, dtype=np.int64)
            b_cnt = np.zeros(nq, dtype=np.int64)
            f_ids = np.full((nq, 2 * min(k, n)), -1, dtype=np.int64)
            f_sc = np.zeros((nq, 2 * min(k, n)), dtype=np.float64)
            f_cnt = np.zeros(nq, dtype=np.int64)
            for q in range(nq):
                res = se.similarity_search_with_embedding(
                    case["queries"][q], df, "voyage-3-large", k, flt)
                rows = [row_of[c] for c in res["id"].tolist()] if len(res) else []
                d_cnt[q] = len(rows)
                d_ids[q, :len(rows)] = rows
                if len(rows):
                    d_sc[q, :len(rows)] = res["similarity"].to_numpy()
                toks = synth.token_strings(case["term_queries"][q])
                hits = se.bm25_search_preprocessed(toks, r_bm25, r_sections, r_section_ids, k, flt)
                docs = [row_of[c] for c in hits]
                b_cnt[q] = len(docs)
                b_ids[q, :len(docs)] = docs
                fused = se.weighted_reciprocal_rank_fusion(
                    [(res["id"].tolist() if len(res) else [], "voyage-3-large"), (hits, "BM25")],
                    WEIGHTS, WRRF_K)
                f_cnt[q] = len(fused)
                f_ids[q, :len(fused)] = [row_of[c] for c, _ in fused]
                f_sc[q, :len(fused)] = [s for _, s in fused]
            out.update({f"dense_ids_{tag}_k{k}": d_ids, f"dense_scores_{tag}_k{k}": d_sc,
                        f"dense_counts_{tag}_k{k}": d_cnt, f"bm25_ids_{tag}_k{k}": b_ids,
                        f"bm25_counts_{tag}_k{k}": b_cnt, f"fused_ids_{tag}_k{k}": f_ids,
                        f"fused_scores_{tag}_k{k}": f_sc, f"fused_counts_{tag}_k{k}": f_cnt})
    # BM25 raw scores of the first queries (float64), straight from get_scores
    out["bm25_all_scores"] = np.stack(
        [r_bm25.get_scores(synth.token_strings(case["term_queries"][q])) for q in range(nq)])
    out["idf_vocab"] = np.array(list(r_bm25.idf.keys()), dtype=object)
    out["idf_values"] = np.array(list(r_bm25.idf.values()), dtype=np.float64)
    out["avgdl"] = np.float64(r_bm25.avgdl)
    return out


def config0_inputs():
    """BASELINE.json configs[0]: 20k x 1024, V=50k, 100 queries x 8 terms (regenerated from
    seeds by the tests; only the reference's outputs are stored)."""
    n, d, vocab = 20000, 1024, 50000
    emb = synth.unit_vectors(n, d, seed=1234)
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=2024)
    queries = synth.unit_vectors(100, d, seed=4321)
    tq = synth.zipf_queries(100, 8, vocab, 1.1, seed=2025)
    srcs = synth.sources(n, seed=77)
    return dict(emb=emb, doc_ptr=doc_ptr, tokens=tokens, queries=queries, term_queries=tq,
                sources=np.array(srcs, dtype=object), vocab=np.int64(vocab))


def checksum(case: dict) -> np.ndarray:
    return np.array([float(case["emb"].astype(np.float64).sum()), float(case["tokens"].sum()),
                     float(case["queries"].astype(np.float64).sum()),
                     float(case["term_queries"].sum())])


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        small = small_case_inputs()
        out = run_reference(small, ks=(10, 100, 3000), filters=(None, "CG,NG", "cg", "ZZ"),
                            tmpdir=tmp)
        np.savez_compressed(os.path.join(GOLDEN, "small_case.npz"), **small, **out)
        print("small_case.npz written")
    if "--no-config0" not in sys.argv:
        with tempfile.TemporaryDirectory() as tmp:
            case = config0_inputs()
            out = run_reference(case, ks=(10,), filters=(None, "CG, NG"), tmpdir=tmp)
            out.pop("bm25_all_scores")
            np.savez_compressed(os.path.join(GOLDEN, "config0_outputs.npz"),
                                input_checksum=checksum(case), **out)
            print("config0_outputs.npz written")


if __name__ == "__main__":
    main()
