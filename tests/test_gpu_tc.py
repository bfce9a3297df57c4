"""GPU: the tensor-core (tcgen05 tf32 + exact fp32 rescoring) scan used for query batches > 8
must return what the exact scan returns: same ids up to oracle-score ties, fp32-accurate scores."""
import importlib
import os

import numpy as np
import pytest

from helpers import check_topk, synth
from oracle import retrieval

pytestmark = pytest.mark.gpu
engine = importlib.import_module("a-nice-rag_b200.engine")


@pytest.fixture(scope="module")
def corpus():
    emb = synth.unit_vectors(20_000, 1024, seed=1234)
    return emb, engine.DenseIndex(emb)


@pytest.mark.parametrize("nq", [9, 32, 33, 64, 70])
@pytest.mark.parametrize("k", [1, 10, 26])
def test_tc_batches_vs_oracle(corpus, nq, k):
    emb, index = corpus
    queries = synth.unit_vectors(nq, 1024, seed=4321 + nq)
    scores, rows, counts = index.search(queries, k)
    full = queries @ emb.T
    for q in range(nq):
        assert counts[q] == k
        want_rows, want_scores = retrieval.dense_topk(queries[q], emb, k)
        check_topk(rows[q], scores[q], want_rows, want_scores, full[q], f"tc nq{nq} k{k} q{q}")


def test_tc_with_mask_and_unnormalised_rows():
    rng = np.random.default_rng(7)
    emb = (synth.unit_vectors(9_000, 512, seed=3) * rng.uniform(0.2, 6.0, size=(9_000, 1))).astype(np.float32)
    index = engine.DenseIndex(emb)
    mask = rng.random(9_000) < 0.6
    queries = (synth.unit_vectors(40, 512, seed=4) * 3.0).astype(np.float32)
    scores, rows, counts = index.search(queries, 10, row_mask=engine.pack_mask(mask))
    for q in range(40):
        want_rows, want_scores = retrieval.dense_topk(queries[q], emb, 10, mask)
        check_topk(rows[q], scores[q], want_rows, want_scores, queries[q] @ emb.T, f"tc mask q{q}")
        assert mask[rows[q]].all()


@pytest.mark.parametrize("nq", [16, 48])
def test_tc_duplicate_cluster_takes_the_exact_fallback(nq):
    """60 identical adjacent rows overflow one warp's candidate list: the query must be flagged
    and rerun through the exact scan, whose tie rule (lower row first) decides."""
    emb = synth.unit_vectors(30_000, 1024, seed=11)
    emb[5000:5060] = emb[5000]
    index = engine.DenseIndex(emb)
    queries = synth.unit_vectors(nq, 1024, seed=12)
    queries[5] = emb[5000]
    queries[nq - 3] = emb[5001]
    scores, rows, counts = index.search(queries, 10)
    assert rows[5].tolist() == list(range(5000, 5010))
    assert rows[nq - 3].tolist() == list(range(5000, 5010))
    for q in range(nq):
        want_rows, want_scores = retrieval.dense_topk(queries[q], emb, 10)
        check_topk(rows[q], scores[q], want_rows, want_scores, queries[q] @ emb.T, f"tc dup q{q}")


def test_tc_pair_and_single_pass_agree(corpus):
    emb, index = corpus
    queries = synth.unit_vectors(64, 1024, seed=77)
    s_pair, r_pair, _ = index.search(queries, 10)
    os.environ["ANR_DISABLE_TC_PAIR"] = "1"
    try:
        s_one, r_one, _ = index.search(queries, 10)
    finally:
        del os.environ["ANR_DISABLE_TC_PAIR"]
    assert np.array_equal(r_pair, r_one) and np.array_equal(s_pair, s_one)


def test_tc_and_scan_paths_agree(corpus):
    emb, index = corpus
    queries = synth.unit_vectors(48, 1024, seed=99)
    s_tc, r_tc, _ = index.search(queries, 10)
    os.environ["ANR_DISABLE_TC"] = "1"
    try:
        s_sc, r_sc, _ = index.search(queries, 10)
    finally:
        del os.environ["ANR_DISABLE_TC"]
    full = queries @ emb.T
    for q in range(48):
        check_topk(r_tc[q], s_tc[q], r_sc[q], s_sc[q], full[q], f"tc vs scan q{q}")
    assert (r_tc == r_sc).mean() > 0.99


@pytest.mark.parametrize("k", [50, 100, 128])
def test_tc_large_k_on_200k_rows(k):
    """top-100-style requests (BASELINE configs[2]) through the tensor-core path: the sample
    pre-pass scales with k, lists stay short, rescoring ranks on exact fp32 scores."""
    import torch
    n, d, nq = 200_000, 1024, 40
    g = torch.Generator(device="cuda").manual_seed(99)
    emb_dev = torch.randn(n, d, generator=g, device="cuda", dtype=torch.float32)
    emb_dev /= emb_dev.norm(dim=1, keepdim=True)
    index = engine.DenseIndex(emb_dev, borrow=True)
    emb = emb_dev.cpu().numpy()
    queries = synth.unit_vectors(nq, d, seed=5)
    scores, rows, counts = index.search(queries, k)
    full = queries @ emb.T
    for q in range(nq):
        assert counts[q] == k
        want_rows, want_scores = retrieval.dense_topk(queries[q], emb, k)
        check_topk(rows[q], scores[q], want_rows, want_scores, full[q], f"tc k{k} q{q}")


# ---- tiled GEMM path (anr_dense_gemm.cu): batches of more than 32 queries on >= ~76k rows ------
@pytest.fixture(scope="module")
def big_corpus():
    import torch
    n, d = 160_000, 1024
    g = torch.Generator(device="cuda").manual_seed(4242)
    emb_dev = torch.randn(n, d, generator=g, device="cuda", dtype=torch.float32)
    emb_dev /= emb_dev.norm(dim=1, keepdim=True)
    return emb_dev.cpu().numpy(), emb_dev


def _check_batch(emb, queries, scores, rows, counts, k, tag, mask=None, every=1):
    for q in range(0, queries.shape[0], every):
        assert counts[q] == k, (tag, q, counts[q])
        want_rows, want_scores = retrieval.dense_topk(queries[q], emb, k, mask)
        check_topk(rows[q], scores[q], want_rows, want_scores, queries[q] @ emb.T, f"{tag} q{q}")


@pytest.mark.parametrize("shadow", [False, True])
@pytest.mark.parametrize("nq,k", [(33, 10), (64, 10), (65, 1), (128, 26), (200, 100), (300, 10)])
def test_gemm_batches_vs_oracle(big_corpus, nq, k, shadow):
    emb, emb_dev = big_corpus
    index = engine.DenseIndex(emb_dev, borrow=True)
    index.set_shadow(shadow)
    queries = synth.unit_vectors(nq, 1024, seed=900 + nq)
    scores, rows, counts = index.search(queries, k)
    _check_batch(emb, queries, scores, rows, counts, k, f"gemm nq{nq} k{k} bf16={shadow}",
                 every=max(nq // 24, 1))


@pytest.mark.parametrize("nq,k", [(1, 10), (3, 25), (9, 10), (32, 100)])
def test_gemm_small_batches_take_the_shadow_pass(big_corpus, nq, k):
    """With a bf16 shadow enabled every batch size goes through the GEMM pass (half the bytes of
    an fp32 scan); results must still be the exact fp32 ranking."""
    emb, emb_dev = big_corpus
    index = engine.DenseIndex(emb_dev, borrow=True)
    index.set_shadow(True)
    queries = synth.unit_vectors(nq, 1024, seed=700 + nq)
    scores, rows, counts = index.search(queries, k)
    _check_batch(emb, queries, scores, rows, counts, k, f"gemm small nq{nq} k{k}")
    index.set_shadow(False)
    s2, r2, _ = index.search(queries, k)
    np.testing.assert_allclose(scores, s2, rtol=1e-5, atol=1e-6)
    assert (rows == r2).mean() > 0.99


@pytest.mark.parametrize("shadow", [False, True])
def test_gemm_with_mask_and_unnormalised_rows(shadow):
    rng = np.random.default_rng(17)
    n, d, nq = 100_000, 512, 70
    emb = (synth.unit_vectors(n, d, seed=33) * rng.uniform(0.2, 6.0, size=(n, 1))).astype(np.float32)
    index = engine.DenseIndex(emb)
    index.set_shadow(shadow)
    mask = rng.random(n) < 0.6
    queries = (synth.unit_vectors(nq, d, seed=34) * 3.0).astype(np.float32)
    scores, rows, counts = index.search(queries, 10, row_mask=engine.pack_mask(mask))
    _check_batch(emb, queries, scores, rows, counts, 10, f"gemm mask bf16={shadow}", mask=mask, every=3)
    assert mask[rows].all()


def test_gemm_candidate_overflow_takes_the_exact_fallback(big_corpus):
    """20 000 identical rows overflow a query's append buffer (16 384 keys): the query is flagged
    and rerun through the exact scan, whose tie rule (lower row first) decides."""
    emb, _ = big_corpus
    emb = emb[:120_000].copy()
    emb[5000:25000] = emb[5000]
    index = engine.DenseIndex(emb)
    queries = synth.unit_vectors(64, 1024, seed=12)
    queries[5] = emb[5000]
    queries[61] = emb[7000]
    scores, rows, counts = index.search(queries, 10)
    assert rows[5].tolist() == list(range(5000, 5010))
    assert rows[61].tolist() == list(range(5000, 5010))
    _check_batch(emb, queries, scores, rows, counts, 10, "gemm overflow", every=7)


def test_gemm_and_scan_paths_agree(big_corpus):
    emb, emb_dev = big_corpus
    index = engine.DenseIndex(emb_dev, borrow=True)
    queries = synth.unit_vectors(96, 1024, seed=99)
    s_g, r_g, _ = index.search(queries, 10)
    index.set_shadow(True)
    s_b, r_b, _ = index.search(queries, 10)
    os.environ["ANR_DISABLE_TC"] = "1"
    try:
        s_sc, r_sc, _ = index.search(queries, 10)
    finally:
        del os.environ["ANR_DISABLE_TC"]
    assert (r_g == r_sc).mean() > 0.99 and (r_b == r_sc).mean() > 0.99
    np.testing.assert_allclose(s_g, s_sc, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s_b, s_sc, rtol=1e-5, atol=1e-6)


def test_hybrid_step_with_reruns_on_both_chains_equals_the_separate_searches(big_corpus):
    """One hybrid call whose dense chain AND BM25 chain both end in a device-side rerun (a query
    whose append buffer overflows on 20 000 duplicate rows; a 60-term query and a query with fewer
    than k matching documents) must return exactly what the dense-only call, the BM25-only call
    and anr_wrrf_fuse return one after the other: the kernels of a chain are launched
    programmatically (each may become resident while its predecessor runs), and the fusion waits
    for the LAST kernel of both chains (search_engine.py:21-34 on query_rag_retrieval.py:357-362's
    two lists).  Repeated, so that a kernel started too early would show."""
    from oracle import csr
    emb, _ = big_corpus
    n, k, nq = 120_000, 10, 64
    emb = emb[:n].copy()
    emb[5000:25000] = emb[5000]
    dense = engine.DenseIndex(emb)
    dense.set_shadow(True)
    vocab = 4000
    ix = csr.from_token_ids(*synth.zipf_corpus(n, vocab, 1.1, seed=51, len_lo=40, len_hi=120),
                            vocab, 1.7, 0.83, 0.05)
    bm25 = engine.Bm25Index(ix.term_ptr, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.k1, ix.b,
                            ix.avgdl)
    queries = synth.unit_vectors(nq, 1024, seed=12)
    queries[5] = emb[5000]
    queries[61] = emb[7000]
    df = np.diff(ix.term_ptr)
    rare = [int(t) for t in np.argsort(np.where(df > 0, df, 1 << 30), kind="stable")[:2]]
    terms = [list(map(int, t)) for t in synth.zipf_queries(nq, 8, vocab, 1.1, seed=77)]
    terms[3] = rare                                                    # fewer than k matches
    terms[40] = [int(t) for t in synth.zipf_queries(1, 60, vocab, 1.1, seed=5)[0]]   # > 48 terms
    terms[41] = []                                                     # no terms: dense list alone
    d_scores, d_rows, d_counts = dense.search(queries, k)
    b_scores, b_docs, b_counts = bm25.search(terms, k)
    assert d_rows[5].tolist() == list(range(5000, 5010))
    ctx = engine.context(dense.ctx_device)
    for rep in range(6):
        got = engine.hybrid_search(dense, bm25, queries, terms, k, k, 5.0, 1.0, 40.0, k,
                                   want_lists=True)
        d_rr, b_rr = ctx.last_rerun()
        assert d_rr >= 2 and b_rr >= 1, (d_rr, b_rr)
        assert np.array_equal(got["dense_rows"], d_rows), rep
        assert np.array_equal(got["dense_scores"], d_scores), rep
        for q in range(nq):
            nb = int(b_counts[q])
            assert np.array_equal(got["bm25_ids"][q, :nb], b_docs[q, :nb]), (rep, q)
            assert np.array_equal(got["bm25_scores"][q, :nb], b_scores[q, :nb]), (rep, q)
            ids, sc = engine.wrrf_fuse([d_rows[q, :int(d_counts[q])].tolist(), b_docs[q, :nb].tolist()],
                                       [5.0, 1.0], 40.0, k)
            c = int(got["counts"][q])
            assert c == len(ids), (rep, q)
            assert np.array_equal(got["ids"][q, :c], ids), (rep, q)
            assert np.array_equal(got["scores"][q, :c], sc), (rep, q)
