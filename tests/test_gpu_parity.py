"""GPU parity tests proper: the CUDA path, called through the C ABI, against the golden
vectors the reference produced and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): dense / BM25 scores within 1e-5 relative; top-k ids
identical except inside groups of (near-)equal oracle score; fused float64 scores and order
bit-identical given the ranked lists."""
import importlib
import os

import numpy as np
import pytest

from helpers import (WEIGHTS, WRRF_K, check_ids_only, check_topk, csr_from_case, filter_mask,
                     synth, tag)
from oracle import csr, pipeline, retrieval

pytestmark = pytest.mark.gpu

engine = importlib.import_module("a-nice-rag_b200.engine")
native = importlib.import_module("a-nice-rag_b200.native")

FILTERS = (None, "CG,NG", "cg")
KS = (10, 100, 3000)


def term_ids_of(case, ix, q):
    return [int(t) if t < ix.idf.shape[0] else -1 for t in case["term_queries"][q]]


@pytest.fixture(scope="module")
def small(small_case):
    ix, okapi = csr_from_case(small_case)
    dense = engine.DenseIndex(small_case["emb"])
    bm25 = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1, ix.b,
                            ix.avgdl)
    return dict(case=small_case, ix=ix, okapi=okapi, dense=dense, bm25=bm25)


# ---------------------------------------------------------------------------------------
# dense
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("flt", FILTERS)
@pytest.mark.parametrize("k", KS)
def test_dense_vs_golden(small, flt, k):
    case = small["case"]
    n = case["emb"].shape[0]
    mask = None if flt is None else filter_mask(case, flt)
    words = None if mask is None else engine.pack_mask(mask)
    scores, rows, counts = small["dense"].search(case["queries"], k, row_mask=words)
    for q in range(case["queries"].shape[0]):
        want_n = int(case[f"dense_counts_{tag(flt)}_k{k}"][q])
        assert counts[q] == want_n == min(k, n if mask is None else int(mask.sum()))
        assert (rows[q, want_n:] == -1).all()
        full = retrieval.dense_scores(case["queries"][q], case["emb"])
        check_topk(rows[q, :want_n], scores[q, :want_n],
                   case[f"dense_ids_{tag(flt)}_k{k}"][q, :want_n],
                   case[f"dense_scores_{tag(flt)}_k{k}"][q, :want_n], full,
                   f"dense q{q} {flt} k{k}")
        if mask is not None:
            assert mask[rows[q, :want_n]].all()


def test_dense_empty_filter_and_single_query(small):
    case = small["case"]
    words = engine.pack_mask(np.zeros(case["emb"].shape[0], dtype=bool))
    scores, rows, counts = small["dense"].search(case["queries"][0], 10, row_mask=words)
    assert counts[0] == 0 and (rows == -1).all() and (scores == 0).all()


@pytest.mark.parametrize("n,d", [(1, 8), (5, 50), (33, 4), (4097, 52), (3000, 1024), (700, 2048),
                                 (257, 3072), (64, 8192)])
@pytest.mark.parametrize("nq", [1, 2, 3, 8, 9])
def test_dense_shapes_vs_oracle(n, d, nq):
    emb = synth.unit_vectors(n, d, seed=100 + n)
    queries = synth.unit_vectors(nq, d, seed=200 + d)
    index = engine.DenseIndex(emb)
    for k in (1, 7, 128, 129):
        scores, rows, counts = index.search(queries, k)
        for q in range(nq):
            c = min(k, n)
            assert counts[q] == c
            want_rows, want_scores = retrieval.dense_topk(queries[q], emb, k)
            check_topk(rows[q, :c], scores[q, :c], want_rows, want_scores,
                       retrieval.dense_scores(queries[q], emb), f"dense n{n} d{d} q{q} k{k}")


def test_dense_exact_ties_prefer_lower_row():
    """All-equal scores: the documented tie rule (higher score, then lower row) decides."""
    emb = np.tile(synth.unit_vectors(1, 64, seed=5), (1000, 1))
    index = engine.DenseIndex(emb)
    scores, rows, counts = index.search(emb[0], 16)
    assert rows[0].tolist() == list(range(16)) and counts[0] == 16
    assert np.all(scores[0] == scores[0, 0])


def test_dense_upload_in_slabs_and_id_base():
    emb = synth.unit_vectors(5000, 128, seed=9)
    index = engine.DenseIndex(n=5000, d=128)
    for r0 in range(0, 5000, 1024):
        index.upload(r0, emb[r0:r0 + 1024])
    q = synth.unit_vectors(1, 128, seed=10)
    scores, rows, _ = index.search(q, 10, id_base=1_000_000)
    want_rows, want_scores = retrieval.dense_topk(q[0], emb, 10)
    check_topk(rows[0] - 1_000_000, scores[0], want_rows, want_scores,
               retrieval.dense_scores(q[0], emb), "slab upload")


def test_dense_200k_rows_property():
    """A size the fused scan runs many tiles per CTA on: compare with the slab-wise oracle."""
    import torch
    n, d = 200_000, 1024
    g = torch.Generator(device="cuda").manual_seed(1234)
    emb = torch.randn(n, d, generator=g, device="cuda", dtype=torch.float32)
    emb /= emb.norm(dim=1, keepdim=True)
    index = engine.DenseIndex(emb, borrow=True)
    queries = synth.unit_vectors(3, d, seed=4321)
    scores, rows, counts = index.search(queries, 10)
    host = emb.cpu().numpy()
    for q in range(3):
        want_rows, want_scores = retrieval.dense_topk(queries[q], host, 10)
        check_topk(rows[q], scores[q], want_rows, want_scores,
                   retrieval.dense_scores(queries[q], host), f"200k q{q}")


# ---------------------------------------------------------------------------------------
# BM25
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("flt", FILTERS)
@pytest.mark.parametrize("k", KS)
def test_bm25_vs_golden(small, flt, k):
    case, ix = small["case"], small["ix"]
    nq = case["term_queries"].shape[0]
    mask = None if flt is None else filter_mask(case, flt)
    words = None if mask is None else engine.pack_mask(mask)
    queries = [term_ids_of(case, ix, q) for q in range(nq)]
    scores, docs, counts = small["bm25"].search(queries, k, doc_mask=words)
    for q in range(nq):
        want_n = int(case[f"bm25_counts_{tag(flt)}_k{k}"][q])
        assert counts[q] == want_n
        want = case[f"bm25_ids_{tag(flt)}_k{k}"][q, :want_n]
        all_scores = case["bm25_all_scores"][q]
        got = docs[q, :want_n]
        if flt:
            # the reference's filtered branch is a stable sort: ascending doc index among exact
            # ties -- the kernel's tie rule; near-ties (fp32 vs float64) stay tolerance-checked
            check_ids_only(got, want, all_scores, f"bm25 q{q} {flt} k{k}")
            assert mask[got].all()
        else:
            check_ids_only(got, want, all_scores, f"bm25 q{q} k{k}")
        np.testing.assert_allclose(scores[q, :want_n], all_scores[got], rtol=1e-5, atol=1e-6)


def test_bm25_filtered_exact_ties_are_stable(small):
    """Planted duplicate documents 39/40/41/1500: equal scores must come out in ascending doc
    order, as the reference's stable sort leaves them (search_engine.py:233)."""
    case, ix = small["case"], small["ix"]
    mask = np.ones(case["emb"].shape[0], dtype=bool)
    scores, docs, counts = small["bm25"].search([term_ids_of(case, ix, 6)], 50,
                                                doc_mask=engine.pack_mask(mask))
    got = docs[0, :counts[0]].tolist()
    pos = [got.index(d) for d in (39, 40, 41, 1500)]
    assert pos == sorted(pos) and pos[-1] - pos[0] == 3
    assert len({float(scores[0, p]) for p in pos}) == 1


def test_bm25_get_scores_vs_oracle(small):
    case, ix = small["case"], small["ix"]
    for q in range(case["term_queries"].shape[0]):
        got = small["bm25"].scores(term_ids_of(case, ix, q))
        np.testing.assert_allclose(got, case["bm25_all_scores"][q], rtol=1e-5, atol=1e-6)


def test_bm25_long_query_and_many_tiles():
    """70 terms (> one staged chunk of 64) over 60k documents (many tiles)."""
    n, vocab = 60_000, 5000
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=31, len_lo=30, len_hi=80)
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                             ix.b, ix.avgdl)
    tq = synth.zipf_queries(5, 70, vocab, 1.1, seed=32)
    for k in (10, 128, 500):
        scores, docs, counts = index.search([list(map(int, t)) for t in tq], k)
        for q in range(5):
            all_scores = csr.scores(ix, [int(t) for t in tq[q]])
            want = retrieval.bm25_topk(all_scores, k)
            check_ids_only(docs[q, :counts[q]], want, all_scores, f"bm25 60k q{q} k{k}")
            np.testing.assert_allclose(scores[q, :counts[q]], all_scores[docs[q, :counts[q]]],
                                       rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------
# fusion
# ---------------------------------------------------------------------------------------
def test_wrrf_bit_exact_random_lists():
    rng = np.random.default_rng(3)
    for trial in range(40):
        n_lists = int(rng.integers(1, 6))
        lists = [rng.permutation(200)[:rng.integers(0, 60)].tolist() for _ in range(n_lists)]
        weights = [float(w) for w in rng.choice([0.5, 1.0, 1.0, 5.0, 2.5], size=n_lists)]
        rrf_k = float(rng.choice([40, 50, 60]))
        want = retrieval.weighted_rrf([(l, str(i)) for i, l in enumerate(lists)],
                                      {str(i): w for i, w in enumerate(weights)}, rrf_k)
        ids, scores = engine.wrrf_fuse(lists, weights, rrf_k)
        assert ids.tolist() == [i for i, _ in want], trial
        assert scores.tolist() == [s for _, s in want], trial


def test_wrrf_small_unions_warp_path_bit_exact():
    """Unions of <= 64 entries (the hybrid query's 2 lists x k) take the warp-per-query kernel:
    overlapping lists, duplicates inside a list, empty lists, exact ties, top_n cuts."""
    rng = np.random.default_rng(13)
    for trial in range(60):
        n_lists = int(rng.integers(1, 5))
        stride = int(rng.integers(1, 64 // n_lists + 1))
        lists = [rng.integers(0, 40, size=rng.integers(0, stride + 1)).tolist() for _ in range(n_lists)]
        lists[0] = (lists[0] + [7] * stride)[:stride]           # longest list fixes the stride
        weights = [float(w) for w in rng.choice([0.5, 1.0, 1.0, 5.0], size=n_lists)]
        rrf_k = float(rng.choice([40, 60]))
        want = retrieval.weighted_rrf([(l, str(i)) for i, l in enumerate(lists)],
                                      {str(i): w for i, w in enumerate(weights)}, rrf_k)
        ids, scores = engine.wrrf_fuse(lists, weights, rrf_k)
        assert ids.tolist() == [i for i, _ in want], (trial, lists)
        assert scores.tolist() == [s for _, s in want], trial
        top_n = int(rng.integers(1, 12))
        ids, scores = engine.wrrf_fuse(lists, weights, rrf_k, top_n=top_n)
        assert ids.tolist() == [i for i, _ in want][:top_n], trial


def test_wrrf_equal_weight_ties_keep_insertion_order():
    a, b = [10, 11, 12, 13], [20, 21, 22, 23]
    ids, scores = engine.wrrf_fuse([a, b], [1.0, 1.0], 40.0)
    assert ids.tolist() == [10, 20, 11, 21, 12, 22, 13, 23]
    ids, _ = engine.wrrf_fuse([a, b], [1.0, 1.0], 40.0, top_n=3)
    assert ids.tolist() == [10, 20, 11]


def test_wrrf_large_lists():
    rng = np.random.default_rng(4)
    lists = [rng.permutation(5000)[:4000].tolist(), rng.permutation(5000)[:4000].tolist()]
    want = retrieval.weighted_rrf([(lists[0], "a"), (lists[1], "b")], {"a": 5.0, "b": 1.0}, 40)
    ids, scores = engine.wrrf_fuse(lists, [5.0, 1.0], 40.0)
    assert ids.tolist() == [i for i, _ in want]
    assert scores.tolist() == [s for _, s in want]


# ---------------------------------------------------------------------------------------
# hybrid + sharded merge
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("flt", (None, "CG,NG"))
def test_hybrid_vs_golden(small, flt):
    case, ix = small["case"], small["ix"]
    nq, k = case["queries"].shape[0], 10
    mask = None if flt is None else filter_mask(case, flt)
    words = None if mask is None else engine.pack_mask(mask)
    queries = [term_ids_of(case, ix, q) for q in range(nq)]
    res = engine.hybrid_search(small["dense"], small["bm25"], case["queries"], queries, k, k,
                               5.0, 1.0, WRRF_K, 2 * k, row_mask=words, doc_mask=words,
                               want_lists=True)
    for q in range(nq):
        full = retrieval.dense_scores(case["queries"][q], case["emb"])
        dn = int(case[f"dense_counts_{tag(flt)}_k{k}"][q])
        check_topk(res["dense_rows"][q, :dn], res["dense_scores"][q, :dn],
                   case[f"dense_ids_{tag(flt)}_k{k}"][q, :dn],
                   case[f"dense_scores_{tag(flt)}_k{k}"][q, :dn], full, f"hybrid dense q{q}")
        bn = int(case[f"bm25_counts_{tag(flt)}_k{k}"][q])
        check_ids_only(res["bm25_ids"][q, :bn], case[f"bm25_ids_{tag(flt)}_k{k}"][q, :bn],
                       case["bm25_all_scores"][q], f"hybrid bm25 q{q}")
        c = int(res["counts"][q])
        pipeline.check_fused(res["ids"][q, :c], res["scores"][q, :c], res["dense_rows"][q, :dn],
                             res["bm25_ids"][q, :bn], (5.0, 1.0), WRRF_K, 2 * k)
        # and against the reference's fused list wherever its two lists had no tie choice
        if (res["dense_rows"][q, :dn] == case[f"dense_ids_{tag(flt)}_k{k}"][q, :dn]).all() and \
                (res["bm25_ids"][q, :bn] == case[f"bm25_ids_{tag(flt)}_k{k}"][q, :bn]).all():
            fn = int(case[f"fused_counts_{tag(flt)}_k{k}"][q])
            assert res["ids"][q, :c].tolist() == case[f"fused_ids_{tag(flt)}_k{k}"][q, :fn].tolist()
            assert res["scores"][q, :c].tolist() == \
                case[f"fused_scores_{tag(flt)}_k{k}"][q, :fn].tolist()


def test_sharded_keys_merge_equals_unsharded(small):
    """3 row shards searched separately, keys concatenated as an all-gather would, merged."""
    case = small["case"]
    emb, queries = case["emb"], case["queries"]
    nq, k = queries.shape[0], 10
    bounds = [0, 700, 1400, emb.shape[0]]
    ctx = engine.context()
    keys = np.zeros((3, nq, k), dtype=np.uint64)
    for s in range(3):
        shard = engine.DenseIndex(emb[bounds[s]:bounds[s + 1]])
        native.call("anr_dense_search_keys", ctx.handle, shard.handle, native.ptr(queries), nq, k,
                    None, bounds[s], native.ptr(keys[s]), None)
    scores = np.empty((nq, k), dtype=np.float32)
    ids = np.empty((nq, k), dtype=np.int32)
    counts = np.empty(nq, dtype=np.int32)
    native.call("anr_topk_merge", ctx.handle, native.ptr(keys), 3, nq, k, native.ptr(scores),
                native.ptr(ids), native.ptr(counts), None)
    u_scores, u_rows, _ = small["dense"].search(queries, k)
    assert np.array_equal(ids, u_rows) and np.array_equal(scores, u_scores)
    assert (counts == k).all()


# ---------------------------------------------------------------------------------------
# the drop-in classes on the reference's own CPU-runnable case (BASELINE configs[0])
# ---------------------------------------------------------------------------------------
def test_dropin_config0_vs_reference_golden(config0_golden, tmp_path):
    from oracle import make_golden
    pkg = importlib.import_module("a-nice-rag_b200")
    case = make_golden.config0_inputs()
    assert np.allclose(make_golden.checksum(case), config0_golden["input_checksum"], rtol=1e-12)
    n = case["emb"].shape[0]
    srcs = list(case["sources"])
    ids = synth.chunk_ids(n, srcs)
    contents = [f"content {i}" for i in range(n)]
    ix, okapi = csr_from_case(case)
    db, pkl = str(tmp_path / "chunks.db"), str(tmp_path / "bm25.pkl")
    synth.write_chunks_db(db, ids, contents, srcs, case["emb"])
    synth.write_bm25_pickle(pkl, okapi, contents, ids, srcs)

    dm = pkg.DatabaseManager()
    df = dm.load_embeddings_from_sql(db, "voyage-3-large")
    bm25, sections, section_ids = dm.load_bm25_from_pickle(pkl)
    assert list(df.columns) == ["id", "document", "source", "embedding", "url"]
    assert dm.load_embeddings_from_sql(db, "voyage-3-large") is df
    se = pkg.SearchEngine(None, None)
    row_of = {c: i for i, c in enumerate(ids)}
    for flt, t in ((None, "none"), ("CG, NG", "CG_NG")):
        for q in range(0, 100, 3):
            res = se.similarity_search_with_embedding(case["queries"][q], df, "voyage-3-large", 10, flt)
            full = retrieval.dense_scores(case["queries"][q], case["emb"])
            check_topk([row_of[c] for c in res["id"]], res["similarity"].to_numpy(),
                       config0_golden[f"dense_ids_{t}_k10"][q], config0_golden[f"dense_scores_{t}_k10"][q],
                       full, f"dropin dense q{q} {flt}")
            assert list(res.columns) == ["id", "document", "source", "embedding", "url", "similarity"]
            toks = synth.token_strings(case["term_queries"][q])
            hits = se.bm25_search_preprocessed(toks, bm25, sections, section_ids, 10, flt)
            all_scores = csr.scores(ix, [int(x) for x in case["term_queries"][q]])
            check_ids_only([row_of[c] for c in hits], config0_golden[f"bm25_ids_{t}_k10"][q],
                           all_scores, f"dropin bm25 q{q} {flt}")
            fused = se.weighted_reciprocal_rank_fusion(
                [(res["id"].tolist(), "voyage-3-large"), (hits, "BM25")], WEIGHTS, WRRF_K)
            want = retrieval.weighted_rrf(
                [(res["id"].tolist(), "voyage-3-large"), (hits, "BM25")], WEIGHTS, WRRF_K)
            assert fused == want
            if q % 9 == 0:
                # the two text-taking wrappers (search_engine.py:100-146 with the caller's vector,
                # the reshape(1, -1) branch :121-123 from a 1-D and from a (1, D) array; :245-269
                # through the tokeniser, with and without lemmatisation: "t123"-style tokens are
                # fixed points of both) against the same goldens
                for vec in (case["queries"][q], case["queries"][q].reshape(1, -1)):
                    res2 = se.similarity_search("ignored text", df, "voyage-3-large", 10, flt,
                                                query_embedding=vec)
                    assert res2["id"].tolist() == res["id"].tolist()
                    assert np.array_equal(res2["similarity"].to_numpy(), res["similarity"].to_numpy())
                    assert list(res2.columns) == list(res.columns)
                text = " ".join(toks).upper() + " ?"
                for lemmatised in (False, True):
                    hits2 = se.bm25_search(text, bm25, sections, section_ids, 10, flt,
                                           use_lemmatized=lemmatised)
                    check_ids_only([row_of[c] for c in hits2],
                                   config0_golden[f"bm25_ids_{t}_k10"][q], all_scores,
                                   f"dropin bm25_search(text) q{q} {flt}")
                    assert hits2 == hits
    # without a vector and without a Voyage client the text search logs and returns an empty frame
    # (search_engine.py:125 -> :148-159 raises, :144-146 returns pd.DataFrame())
    assert se.similarity_search("some text", df, "voyage-3-large", 10).empty
    assert se.bm25_search("", bm25, sections, section_ids, 10) == []
    assert se.bm25_search("the of and 12", bm25, sections, section_ids, 10) == []   # stop-words only
    # batched extension == per-query calls
    qs = list(range(0, 16))
    batch = se.hybrid_search_batch(case["queries"][qs],
                                   [synth.token_strings(case["term_queries"][q]) for q in qs],
                                   df, bm25, sections, section_ids, WEIGHTS, "voyage-3-large",
                                   10, 10, WRRF_K)
    for j, q in enumerate(qs):
        res = se.similarity_search_with_embedding(case["queries"][q], df, "voyage-3-large", 10)
        hits = se.bm25_search_preprocessed(synth.token_strings(case["term_queries"][q]), bm25,
                                           sections, section_ids, 10)
        want = retrieval.weighted_rrf([(res["id"].tolist(), "voyage-3-large"), (hits, "BM25")],
                                      WEIGHTS, WRRF_K)[:10]
        assert batch[j] == want
    # An EMPTY token list has no BM25 list at all (search_engine.py:216-217 returns [] before any
    # scoring: the query is fused from its dense list alone); a list of unknown tokens is scored
    # (all zeros) and does contribute k zero-score documents, as in the reference.
    odd_tokens = [[], ["zzz-not-in-the-vocabulary"], synth.token_strings(case["term_queries"][2])]
    batch = se.hybrid_search_batch(case["queries"][:3], odd_tokens, df, bm25, sections, section_ids,
                                   WEIGHTS, "voyage-3-large", 10, 10, WRRF_K)
    for j, toks in enumerate(odd_tokens):
        res = se.similarity_search_with_embedding(case["queries"][j], df, "voyage-3-large", 10)
        hits = se.bm25_search_preprocessed(toks, bm25, sections, section_ids, 10)
        assert len(hits) == (0 if not toks else 10)
        want = retrieval.weighted_rrf([(res["id"].tolist(), "voyage-3-large"), (hits, "BM25")],
                                      WEIGHTS, WRRF_K)[:10]
        assert batch[j] == want, j
    assert [i for i, _ in batch[0]] == se.similarity_search_with_embedding(
        case["queries"][0], df, "voyage-3-large", 10)["id"].tolist()
    # the same through the index classes: counts 0 / ids -1 for the query without terms
    b_ix = importlib.import_module("a-nice-rag_b200.registry").resolve_bm25(bm25).index
    sc, dc, ct = b_ix.search([[], [int(x) for x in case["term_queries"][2]], [-1]], 10)
    assert ct.tolist() == [0, 10, 10] and (dc[0] == -1).all() and (sc[0] == 0).all()
    assert (sc[2] == 0).all() and dc[2].tolist() == list(range(10))


def test_wrrf_evaluator_sized_union_bit_exact():
    """5 ranked lists x 12 000 ids (retrieval_eval.py:142-143: similarity_k = 12000): the union
    does not fit in shared memory and runs through the global scratch."""
    rng = np.random.default_rng(8)
    lists = [rng.permutation(15000)[:12000].tolist() for _ in range(5)]
    weights = [5.0, 0.7, 1.3, 2.0, 1.0]
    want = retrieval.weighted_rrf([(l, str(i)) for i, l in enumerate(lists)],
                                  {str(i): w for i, w in enumerate(weights)}, 40)
    ids, scores = engine.wrrf_fuse(lists, weights, 40.0)
    assert ids.tolist() == [i for i, _ in want]
    assert scores.tolist() == [s for _, s in want]


def test_sharded_fuse_equals_unsharded_hybrid(small):
    """Three shards (dense rows AND the same documents' postings, global BM25 statistics) searched
    separately; their keys, laid out as the NCCL all-gather would leave them, go through
    anr_sharded_fuse and must reproduce the unsharded anr_hybrid_search bit for bit."""
    case, ix = small["case"], small["ix"]
    emb, queries = case["emb"], case["queries"]
    nq, k = queries.shape[0], 10
    term_queries = [term_ids_of(case, ix, q) for q in range(nq)]
    terms, offsets = engine.Bm25Index.pack_queries(term_queries)
    bounds = [0, 700, 1400, emb.shape[0]]
    ctx = engine.context()
    gathered = np.zeros((3, 2, nq, k), dtype=np.uint64)
    for s in range(3):
        lo, hi = bounds[s], bounds[s + 1]
        d_shard = engine.DenseIndex(emb[lo:hi])
        local = csr.from_token_ids(case["doc_ptr"][lo:hi + 1] - case["doc_ptr"][lo],
                                   case["tokens"][case["doc_ptr"][lo]:case["doc_ptr"][hi]],
                                   int(case["vocab"]), ix.k1, ix.b, 0.05)
        b_shard = engine.Bm25Index(local.term_ptr, local.post_doc, local.post_tf, local.doc_len,
                                   ix.idf, ix.k1, ix.b, ix.avgdl)      # GLOBAL idf / avgdl
        native.call("anr_dense_search_keys", ctx.handle, d_shard.handle, native.ptr(queries), nq,
                    k, None, lo, native.ptr(gathered[s, 0]), None)
        native.call("anr_bm25_search_keys", ctx.handle, b_shard.handle, native.ptr(terms),
                    native.ptr(offsets), nq, k, None, None, lo, native.ptr(gathered[s, 1]), None)
    ids = np.empty((nq, 2 * k), dtype=np.int32)
    scores = np.empty((nq, 2 * k), dtype=np.float64)
    counts = np.empty(nq, dtype=np.int32)
    native.call("anr_sharded_fuse", ctx.handle, native.ptr(gathered), 3, nq, k, 5.0, 1.0,
                float(WRRF_K), 2 * k, native.ptr(ids), native.ptr(scores), native.ptr(counts), None)
    res = engine.hybrid_search(small["dense"], small["bm25"], queries, term_queries, k, k, 5.0, 1.0,
                               WRRF_K, 2 * k)
    assert np.array_equal(counts, res["counts"])
    assert np.array_equal(ids, res["ids"]) and np.array_equal(scores, res["scores"])


def test_hybrid_search_keys_equals_the_two_key_searches(small):
    """anr_hybrid_search_keys (what a shard calls: both searches in one call, BM25 on the side
    stream around the dense pass) returns exactly the keys of the two separate calls."""
    import torch
    case, ix = small["case"], small["ix"]
    queries = case["queries"]
    nq, k = queries.shape[0], 10
    term_queries = [term_ids_of(case, ix, q) for q in range(nq)]
    terms, offsets = engine.Bm25Index.pack_queries(term_queries)
    ctx = engine.context()
    dev = torch.device("cuda", 0)
    q_dev = torch.from_numpy(np.ascontiguousarray(queries)).to(dev)
    t_dev, o_dev = torch.from_numpy(terms).to(dev), torch.from_numpy(offsets).to(dev)
    both = torch.zeros((2, nq, k), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    native.call("anr_hybrid_search_keys", ctx.handle, small["dense"].handle, small["bm25"].handle,
                q_dev.data_ptr(), t_dev.data_ptr(), o_dev.data_ptr(), nq, k, None, None, 1000, 5000,
                both.data_ptr(), None)
    native.call("anr_ctx_sync", ctx.handle)
    want = np.zeros((2, nq, k), dtype=np.uint64)
    native.call("anr_dense_search_keys", ctx.handle, small["dense"].handle, native.ptr(queries), nq, k,
                None, 1000, native.ptr(want[0]), None)
    native.call("anr_bm25_search_keys", ctx.handle, small["bm25"].handle, native.ptr(terms),
                native.ptr(offsets), nq, k, None, None, 5000, native.ptr(want[1]), None)
    assert np.array_equal(both.cpu().numpy().view(np.uint64), want)


def test_bm25_long_posting_lists_many_tiles():
    """Posting lists of 10^4..10^5 entries cut by many document tiles: exercises every level of
    the 32-ary slice search, including terms whose postings all lie before / after a tile."""
    n, vocab = 300_000, 300
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=41, len_lo=8, len_hi=20)
    tokens[doc_ptr[200_000]:] = np.minimum(tokens[doc_ptr[200_000]:], 40)   # rare terms end early
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                             ix.b, ix.avgdl)
    tq = synth.zipf_queries(6, 8, vocab, 1.1, seed=42)
    tq[0, :] = [299, 250, 200, 150, 100, 60, 45, 41]       # lists that stop at document 200 000
    for k in (10, 100):
        scores, docs, counts = index.search([list(map(int, t)) for t in tq], k)
        for q in range(6):
            all_scores = csr.scores(ix, [int(t) for t in tq[q]])
            want = retrieval.bm25_topk(all_scores, k)
            check_ids_only(docs[q, :counts[q]], want, all_scores, f"bm25 long q{q} k{k}")
            np.testing.assert_allclose(scores[q, :counts[q]], all_scores[docs[q, :counts[q]]],
                                       rtol=1e-5, atol=1e-6)


def test_bm25_reweight_equals_rebuilt_index(small):
    """k1 / b / epsilon sweep step (src/processing/bm25_test.py): re-weighting the resident
    postings must equal building a fresh index with the new parameters."""
    from oracle import bm25_okapi
    case, ix = small["case"], small["ix"]
    corpus = synth.doc_token_lists(case["doc_ptr"], case["tokens"])
    other = bm25_okapi.BM25Okapi(corpus, k1=0.857, b=0.686, epsilon=0.075)   # results/bm25_optimization_results.csv:2
    idf = np.zeros(int(case["vocab"]), dtype=np.float64)
    for tok, val in other.idf.items():
        idf[int(tok[1:])] = val
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1, ix.b,
                             ix.avgdl)
    index.reweight(ix.post_tf, ix.doc_len, idf, 0.857, 0.686, other.avgdl)
    for q in range(4):
        toks = synth.token_strings(case["term_queries"][q])
        got = index.scores(term_ids_of(case, ix, q))
        np.testing.assert_allclose(got, other.get_scores(toks), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------
# BM25 top-k on 8192+ documents, both safe-pruning paths: the candidate-driven one
# (anr_bm25_ms.cu, the default) and the tiled scan with dense head rows (anr_bm25.cu, PRUNE;
# ANR_BM25_MAXSCORE=0 selects it, and it is the rerun path of flagged queries)
@pytest.fixture(params=["candidates", "tiled"])
def bm25_path(request, monkeypatch):
    if request.param == "tiled":
        monkeypatch.setenv("ANR_BM25_MAXSCORE", "0")
    return request.param


@pytest.fixture(scope="module")
def prune_corpus():
    n, vocab = 70_000, 4000
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=51, len_lo=40, len_hi=120)
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                             ix.b, ix.avgdl)
    return ix, index, vocab


def _check_bm25_batch(ix, queries, scores, docs, counts, k, what, mask=None):
    n_terms = len(ix.term_ptr) - 1
    for q, terms in enumerate(queries):
        if len(terms) == 0:   # `if not query_tokens: return []` (search_engine.py:216-217)
            assert counts[q] == 0 and (docs[q] == -1).all() and (scores[q] == 0).all(), (what, q)
            continue
        all_scores = csr.scores(ix, [int(t) if 0 <= int(t) < n_terms else -1 for t in terms])
        if mask is None:
            want = retrieval.bm25_topk(all_scores, k)
        else:   # the filtered branch (search_engine.py:224-233): stable descending sort of the kept docs
            kept = np.flatnonzero(mask)
            want = kept[np.argsort(-all_scores[kept], kind="stable")][:k]
        assert counts[q] == len(want), (what, q, counts[q], len(want))
        check_ids_only(docs[q, :counts[q]], want, all_scores, f"{what} q{q} k{k}")
        np.testing.assert_allclose(scores[q, :counts[q]], all_scores[docs[q, :counts[q]]],
                                   rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("nq,k", [(1, 10), (8, 10), (16, 1), (17, 100), (64, 10), (70, 32), (33, 128)])
def test_bm25_pruned_batches_vs_oracle(prune_corpus, bm25_path, nq, k):
    ix, index, vocab = prune_corpus
    tq = synth.zipf_queries(nq, 8, vocab, 1.1, seed=60 + nq)
    queries = [list(map(int, t)) for t in tq]
    if nq > 3:
        queries[1] = queries[1] + queries[1][:3]            # duplicate terms count twice
        queries[2] = [vocab + 5, -1] + queries[2]            # unknown terms contribute nothing
        queries[3] = []                                      # empty query: no result at all
        queries[4] = [vocab + 9, -1]                         # unknown terms only: k zero-score docs
        queries[0] = [0, 1, 2, 3, 0]                         # head terms only: nothing is pruned
    scores, docs, counts = index.search(queries, k)
    _check_bm25_batch(ix, queries, scores, docs, counts, k, f"pruned nq{nq}")


def test_bm25_pruned_agrees_with_unpruned_scan(prune_corpus, bm25_path):
    ix, index, vocab = prune_corpus
    tq = synth.zipf_queries(40, 8, vocab, 1.1, seed=77)
    queries = [list(map(int, t)) for t in tq]
    s_g, d_g, c_g = index.search(queries, 10)
    os.environ["ANR_DISABLE_BM25_PRUNE"] = "1"
    try:
        s_o, d_o, c_o = index.search(queries, 10)
    finally:
        del os.environ["ANR_DISABLE_BM25_PRUNE"]
    assert np.array_equal(c_g, c_o)
    np.testing.assert_allclose(s_g, s_o, rtol=1e-5, atol=1e-6)
    assert (d_g == d_o).mean() > 0.98


def test_bm25_pruned_long_queries_take_the_unpruned_path(prune_corpus, bm25_path):
    """More than 64 terms in a query (one staged chunk): that query is scanned unpruned, next to
    pruned 3- and 33-term queries in the same launch."""
    ix, index, vocab = prune_corpus
    long_q = synth.zipf_queries(20, 70, vocab, 1.1, seed=81)
    queries = [list(map(int, t)) for t in long_q]
    queries[4] = queries[4][:3]
    queries[9] = queries[9][:33]
    scores, docs, counts = index.search(queries, 10)
    _check_bm25_batch(ix, queries, scores, docs, counts, 10, "pruned long")


def test_bm25_pruned_with_doc_mask(prune_corpus, bm25_path):
    ix, index, vocab = prune_corpus
    rng = np.random.default_rng(5)
    mask = rng.random(ix.doc_len.shape[0]) < 0.3
    tq = synth.zipf_queries(40, 8, vocab, 1.1, seed=91)
    queries = [list(map(int, t)) for t in tq]
    scores, docs, counts = index.search(queries, 10, doc_mask=engine.pack_mask(mask))
    _check_bm25_batch(ix, queries, scores, docs, counts, 10, "pruned mask", mask=mask)
    assert mask[docs].all()


def test_bm25_pruned_after_reweight(prune_corpus, bm25_path):
    ix, index, vocab = prune_corpus
    other = csr.from_token_ids(*_prune_tokens(), vocab, 0.9, 0.7, 0.075)
    index2 = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                              ix.b, ix.avgdl)
    tq = synth.zipf_queries(16, 8, vocab, 1.1, seed=95)
    queries = [list(map(int, t)) for t in tq]
    index2.search(queries, 10)                      # builds the dense head rows with the OLD weights
    index2.reweight(other.post_tf, other.doc_len, other.idf, 0.9, 0.7, other.avgdl)
    scores, docs, counts = index2.search(queries, 10)
    _check_bm25_batch(other, queries, scores, docs, counts, 10, "pruned reweight")


def test_bm25_content_word_queries(prune_corpus, bm25_path):
    """Stop-word-free queries (SURVEY 8d's "content-word" variant: terms resampled within Zipf ranks
    > 27, what preprocess_bm25.py:41-46 leaves of a real query): eight terms of similar idf give the
    candidate-driven path a weak bound -- few lists set aside, nearly every posting a candidate --
    which is the distribution where its pruning decisions are exercised most."""
    ix, index, vocab = prune_corpus
    tq = synth.zipf_queries(48, 8, vocab, 1.1, seed=131, skip_head=27)
    queries = [list(map(int, t)) for t in tq]
    for k in (10, 100):
        scores, docs, counts = index.search(queries, k)
        _check_bm25_batch(ix, queries, scores, docs, counts, k, f"content words k{k}")


def _prune_tokens():
    return synth.zipf_corpus(70_000, 4000, 1.1, seed=51, len_lo=40, len_hi=120)


def test_bm25_candidate_path_flag_conditions_and_repeatability(prune_corpus):
    """The candidate-driven path (anr_bm25_ms.cu) next to the cases it hands to the exhaustive
    scan on the device: fewer than k documents with a positive score (the reference ranks
    zero-score documents too, search_engine.py:236-241), more survivors than its buffer holds
    (one very frequent term, k = 128), more than 48 terms, a filter that keeps nothing; and its
    scores do not depend on timing: two calls return identical bits."""
    ix, index, vocab = prune_corpus
    df = np.diff(ix.term_ptr)
    rare = [int(t) for t in np.argsort(np.where(df > 0, df, 1 << 30), kind="stable")[:2]]   # shortest lists
    tq = synth.zipf_queries(24, 8, vocab, 1.1, seed=123)
    queries = [list(map(int, t)) for t in tq]
    queries[0] = rare                         # (next to) no matching documents
    queries[1] = [0]                          # the most frequent term alone
    queries[2] = [int(t) for t in synth.zipf_queries(1, 60, vocab, 1.1, seed=5)[0]]
    queries[3] = []
    queries[4] = [rare[0]] * 3 + [0]          # duplicates of a rare term beside a head term
    ctx = engine.context(index.ctx_device)
    for k in (10, 128):
        a = index.search(queries, k)
        reran = ctx.last_rerun()[1]
        b = index.search(queries, k)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        _check_bm25_batch(ix, queries, *a, k, f"candidates k{k}")
        # the 60-term query always goes through the exhaustive scan; the empty one never does
        assert 1 <= reran < len(queries), reran
    index.search(queries[5:], 10)
    assert ctx.last_rerun() == (-1, 0)        # ordinary queries: nothing left the candidate path
    nothing = np.zeros(ix.doc_len.shape[0], dtype=bool)
    scores, docs, counts = index.search(queries, 10, doc_mask=engine.pack_mask(nothing))
    assert (counts == 0).all() and (docs == -1).all()
    few = nothing.copy()
    few[[5, 77, 40_000]] = True
    scores, docs, counts = index.search(queries, 10, doc_mask=engine.pack_mask(few))
    _check_bm25_batch(ix, queries, scores, docs, counts, 10, "candidates 3 docs kept", mask=few)


def test_bm25_negative_average_idf_takes_the_exhaustive_scan():
    """Every term in more than half of the documents: BM25Okapi's raw idf values are all negative,
    so is their average, and the epsilon floor (epsilon * average_idf, rank_bm25 _calc_idf) is
    NEGATIVE too.  Term bounds mean nothing then: the candidate-driven path must hand the whole
    index to the exhaustive scan, which ranks documents that lack the terms (score 0) first."""
    rng = np.random.default_rng(7)
    n, vocab = 3000, 6
    doc_len = rng.integers(10, 16, n)
    doc_ptr = np.concatenate([[0], np.cumsum(doc_len)]).astype(np.int64)
    tokens = rng.integers(0, vocab, int(doc_ptr[-1])).astype(np.int32)
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.5, 0.75, 0.25)
    assert (ix.idf < 0).all()
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                             ix.b, ix.avgdl)
    queries = [[0, 1], [2], [3, 3, 4], [5, 0, 1, 2]] * 5
    for k in (10, 100):
        scores, docs, counts = index.search(queries, k)
        _check_bm25_batch(ix, queries, scores, docs, counts, k, f"negative idf k{k}")


def test_bm25_candidate_path_ties_duplicates_and_term_count_limits():
    """A corpus of 60 distinct documents, each stored 100 times (every score occurs 100 times: the
    k-th best is always inside a run of equal scores, and far more than k documents reach theta),
    and queries of exactly 48 terms (the most the candidate-driven path takes), 49 (flagged and
    rerun) and 1."""
    rng = np.random.default_rng(11)
    vocab, distinct, copies = 400, 60, 100
    base_docs = [rng.integers(0, vocab, rng.integers(20, 60)).astype(np.int32) for _ in range(distinct)]
    order = rng.permutation(distinct * copies) % distinct
    docs_tok = [base_docs[i] for i in order]
    doc_ptr = np.concatenate([[0], np.cumsum([len(d) for d in docs_tok])]).astype(np.int64)
    tokens = np.concatenate(docs_tok)
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    index = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1,
                             ix.b, ix.avgdl)
    q48 = [int(t) for t in rng.integers(0, vocab, 48)]
    queries = [q48, q48 + [int(q48[0])], [int(base_docs[0][0])],
               [int(t) for t in base_docs[3][:6]], [int(t) for t in base_docs[7][:3]] * 2]
    for k in (1, 10, 100, 128):
        scores, docs, counts = index.search(queries, k)
        _check_bm25_batch(ix, queries, scores, docs, counts, k, f"ties k{k}")
        for q in range(len(queries)):   # one total order everywhere: equal scores -> lower document first
            sc, dd = scores[q, :counts[q]], docs[q, :counts[q]]
            same = sc[1:] == sc[:-1]
            assert (dd[1:][same] > dd[:-1][same]).all(), (k, q)


# ---------------------------------------------------------------------------------------
# batched orchestrator: retrieve_documents for B queries == the per-query pipeline
def test_retrieve_documents_batch_equals_per_query_orchestrator(small):
    """a-nice-rag_b200.batch_retrieval.retrieve_documents_batch vs oracle.orchestrator (the
    restatement of query_rag_retrieval.py:149-411 that tests/test_oracle.py pins against the
    unmodified reference method), both over the drop-in SearchEngine: two dense models + BM25,
    filters, full ranking (similarity_k >= N), return_docs."""
    import types
    import pandas as pd
    from oracle import orchestrator
    pkg = importlib.import_module("a-nice-rag_b200")
    batch = importlib.import_module("a-nice-rag_b200.batch_retrieval")
    case = small["case"]
    n = case["emb"].shape[0]
    srcs = list(case["sources"])
    ids = synth.chunk_ids(n, srcs)
    rng = np.random.default_rng(3)
    emb2 = (case["emb"] + 0.3 * rng.standard_normal(case["emb"].shape)).astype(np.float32)

    def frame(e):
        return pd.DataFrame({"id": ids, "document": [f"doc {i}" for i in range(n)], "source": srcs,
                             "embedding": list(e), "url": [""] * n})
    sections = [types.SimpleNamespace(page_content=f"doc {i}", metadata={"id": ids[i], "source": srcs[i]})
                for i in range(n)]
    nice = pkg.InfoSource("nice")
    system = types.SimpleNamespace(
        config=pkg.Config(), search_engine=pkg.SearchEngine(None, None),
        embeddings_data={nice: {"voyage-3-large": frame(case["emb"]), "Qwen3": frame(emb2)}},
        bm25_data={nice: (small["okapi"], sections, ids)})
    weights = {"voyage-3-large": 5.0, "Qwen3": 2.0, "BM25": 1.0}
    nq = min(12, case["queries"].shape[0])
    qe = {"voyage-3-large": np.ascontiguousarray(case["queries"][:nq]),
          "Qwen3": np.ascontiguousarray(np.roll(case["queries"][:nq], 1, axis=0))}
    toks = [synth.token_strings(case["term_queries"][q]) for q in range(nq)]
    toks[nq // 2] = []                                    # no BM25 list for this query
    cases = [dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True),
             dict(similarity_k=10, common_sections_n=7, use_hybrid_search=False),
             dict(similarity_k=25, common_sections_n=15, use_hybrid_search=True,
                  filename_type_filter="CG, NG"),
             dict(similarity_k=3000, common_sections_n=40, use_hybrid_search=True),
             dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True, return_docs=True),
             dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True,
                  filename_type_filter="ZZ")]
    for kw in cases:
        got = batch.retrieve_documents_batch(system, qe, None, toks, info_source="NICE",
                                             model_weights=weights, wrrf_k=40, **kw)
        assert len(got) == nq
        for q in range(nq):
            one = {m: e[q] for m, e in qe.items()}
            want = orchestrator.retrieve_documents(system, nice, one, None, toks[q],
                                                   model_weights=weights, wrrf_k=40, **kw)
            if kw.get("return_docs"):
                assert [d["id"] for d in got[q]] == [d["id"] for d in want], (kw, q)
                np.testing.assert_allclose([d["similarity"] for d in got[q]],
                                           [d["similarity"] for d in want], rtol=1e-6)
            else:
                assert got[q] == want, (kw, q)


def test_bm25_published_known_answer_on_the_device():
    """rank_bm25's README example (tests/test_oracle.py::test_bm25_okapi_published_known_answer)
    through the product path: pickle-shaped object -> CSR inversion -> anr_bm25_scores / search."""
    from oracle import bm25_okapi
    corpus = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
    okapi = bm25_okapi.BM25Okapi([doc.split(" ") for doc in corpus])
    index = engine.Bm25Index.from_okapi(okapi)
    terms = index.term_ids(["windy", "London"])
    np.testing.assert_allclose(index.scores(terms), [0.0, 0.93729472, 0.0], rtol=1e-6, atol=0)
    scores, docs, counts = index.search([terms], 1)
    assert counts[0] == 1 and docs[0, 0] == 1
    np.testing.assert_allclose(scores[0, 0], 0.93729472, rtol=1e-6)
