"""GPU: BASELINE.json configs[1] at FULL size (1M chunks x 1024-d + BM25 over 1M docs, batch 64,
top-10 WRRF) through the C ABI.  The CPU oracle cannot finish this size in seconds, so parity is
checked through size-independent properties on the same device tensors:

* dense: the returned rows of sampled queries equal torch's fp32 matmul + topk (allow_tf32 off) up to
  score ties, the returned scores are the exact fp32 inner products;
* BM25: the returned documents of sampled queries carry the top-k scores of a float64 scatter of the
  same postings, the returned scores match it within 1e-5 relative;
* fusion: given the two lists the kernels returned, the fused ids / float64 scores are bit-identical
  to the Python statement of weighted RRF (oracle.retrieval.weighted_rrf);
* the bf16-shadow pass, the tf32 pass and a repeated call return identical results (determinism,
  exact rescoring), and batch-1 equals row q of the batch.
"""
import importlib

import numpy as np
import pytest

from oracle import retrieval

pytestmark = pytest.mark.gpu
engine = importlib.import_module("a-nice-rag_b200.engine")
native = importlib.import_module("a-nice-rag_b200.native")
synth = importlib.import_module("a-nice-rag_b200.synth")

N, D, VOCAB, B, K = 1_000_000, 1024, 50_000, 64, 10
W_DENSE, W_BM25, WRRF_K = 5.0, 1.0, 40.0


@pytest.fixture(scope="module")
def config1():
    import torch
    dev = torch.device("cuda", 0)
    emb = synth.unit_vectors_torch(N, D, 1234, dev)
    post = synth.zipf_postings_torch(N, VOCAB, 1.1, 2024, dev)
    idf = synth.idf_from_counts(N, post["nd"].cpu().numpy(), 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / N
    dense = engine.DenseIndex(emb, borrow=True)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                            idf, 1.7, 0.83, avgdl, n_terms=VOCAB, n_docs=N)
    queries = synth.unit_vectors(B, D, seed=4321)
    terms = synth.zipf_queries(B, 8, VOCAB, 1.1, seed=2025)
    return dict(emb=emb, post=post, idf=idf, avgdl=avgdl, dense=dense, bm25=bm25, queries=queries,
                terms=terms)


def _hybrid(c, queries, terms):
    return engine.hybrid_search(c["dense"], c["bm25"], queries, [list(map(int, t)) for t in terms],
                                K, K, W_DENSE, W_BM25, WRRF_K, K, want_lists=True)


def test_config1_hybrid_batch64_properties(config1):
    import torch
    c = config1
    c["dense"].set_shadow(True)
    got = _hybrid(c, c["queries"], c["terms"])
    torch.backends.cuda.matmul.allow_tf32 = False
    q_dev = torch.from_numpy(c["queries"]).to(c["emb"].device)
    tp, pdoc, ptf, dl = (c["post"][k] for k in ("term_ptr", "post_doc", "post_tf", "doc_len"))
    for q in range(0, B, 9):
        # dense
        ref = torch.mv(c["emb"], q_dev[q])
        top = torch.topk(ref, K)
        rows = torch.from_numpy(got["dense_rows"][q].astype(np.int64)).to(ref.device)
        assert torch.allclose(ref[rows], top.values, rtol=1e-5, atol=1e-6), f"dense ranking q{q}"
        np.testing.assert_allclose(got["dense_scores"][q], ref[rows].cpu().numpy(), rtol=1e-5, atol=1e-6)
        assert len(set(got["dense_rows"][q].tolist())) == K
        # BM25
        acc = torch.zeros(N, dtype=torch.float64, device=ref.device)
        for term in c["terms"][q]:
            lo, hi = int(tp[term]), int(tp[term + 1])
            dd = pdoc[lo:hi].long()
            tf = ptf[lo:hi].double()
            acc[dd] += c["idf"][term] * (tf * 2.7 / (tf + 1.7 * (1 - 0.83 + 0.83 * dl[dd].double() / c["avgdl"])))
        btop = torch.topk(acc, K)
        docs = torch.from_numpy(got["bm25_ids"][q].astype(np.int64)).to(ref.device)
        assert torch.allclose(acc[docs], btop.values, rtol=1e-5, atol=1e-6), f"bm25 ranking q{q}"
        np.testing.assert_allclose(got["bm25_scores"][q].astype(np.float64), acc[docs].cpu().numpy(),
                                   rtol=1e-5, atol=1e-6)
    # fusion: exact given the lists, every query
    for q in range(B):
        want = retrieval.weighted_rrf(
            [([int(i) for i in got["dense_rows"][q]], "d"), ([int(i) for i in got["bm25_ids"][q]], "b")],
            {"d": W_DENSE, "b": W_BM25}, WRRF_K)[:K]
        n = int(got["counts"][q])
        assert [int(i) for i in got["ids"][q, :n]] == [i for i, _ in want]
        assert [float(s) for s in got["scores"][q, :n]] == [s for _, s in want]


def test_config1_paths_agree_and_are_deterministic(config1):
    c = config1
    c["dense"].set_shadow(True)
    a = _hybrid(c, c["queries"], c["terms"])
    b = _hybrid(c, c["queries"], c["terms"])
    for key in ("ids", "scores", "counts", "dense_rows", "dense_scores", "bm25_ids", "bm25_scores"):
        assert np.array_equal(a[key], b[key]), f"repeat call differs in {key}"
    c["dense"].set_shadow(False)
    t = _hybrid(c, c["queries"], c["terms"])            # tf32 operands on the fp32 words
    assert np.array_equal(a["dense_rows"], t["dense_rows"])
    assert np.array_equal(a["dense_scores"], t["dense_scores"])      # both are exact fp32 rescoring
    assert np.array_equal(a["ids"], t["ids"])
    one = _hybrid(c, c["queries"][5:6], c["terms"][5:6])            # batch-1: fp32 CUDA-core scan
    np.testing.assert_allclose(one["dense_scores"][0], a["dense_scores"][5], rtol=1e-6, atol=1e-7)
    assert np.array_equal(one["dense_rows"][0], a["dense_rows"][5])
    assert np.array_equal(one["bm25_ids"][0], a["bm25_ids"][5])
    assert np.array_equal(one["ids"][0], a["ids"][5])


def test_config2_10M_batch1024_top100_properties():
    """BASELINE.json configs[2] at full size: 10M x 1024 fp32, batch 1024, dense top-100 through the
    cta_group::2 GEMM pass (tf32 operands, then the bf16 shadow): sampled queries against torch fp32
    matmul + topk on the same device tensor; both operand modes return identical results."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * (1 << 30):
        pytest.skip("needs ~65 GB of device memory")
    n, b, k = 10_000_000, 1024, 100
    dev = torch.device("cuda", 0)
    emb = synth.unit_vectors_torch(n, D, 1234, dev)
    index = engine.DenseIndex(emb, borrow=True)
    queries = synth.unit_vectors(b, D, seed=4321)
    s_tf32, r_tf32, c_tf32 = index.search(queries, k)
    index.set_shadow(True)
    s_bf16, r_bf16, c_bf16 = index.search(queries, k)
    assert (c_tf32 == k).all() and (c_bf16 == k).all()
    assert np.array_equal(r_tf32, r_bf16) and np.array_equal(s_tf32, s_bf16)
    torch.backends.cuda.matmul.allow_tf32 = False
    q_dev = torch.from_numpy(queries).to(dev)
    for q in range(0, b, 171):
        ref = torch.mv(emb, q_dev[q])
        top = torch.topk(ref, k)
        rows = torch.from_numpy(r_bf16[q].astype(np.int64)).to(dev)
        assert torch.allclose(ref[rows], top.values, rtol=1e-5, atol=1e-6), f"ranking q{q}"
        np.testing.assert_allclose(s_bf16[q], ref[rows].cpu().numpy(), rtol=1e-5, atol=1e-6)
        assert len(set(r_bf16[q].tolist())) == k
        assert np.all(np.diff(s_bf16[q]) <= 0)          # best first


def _shard_postings(torch, post, vocab, lo, hi):
    """The postings of documents [lo, hi) as a CSR index of their own (local document ids)."""
    dev = post["post_doc"].device
    term_of = torch.repeat_interleave(torch.arange(vocab, device=dev), post["nd"])
    keep = (post["post_doc"] >= lo) & (post["post_doc"] < hi)
    nd = torch.bincount(term_of[keep], minlength=vocab)
    term_ptr = torch.zeros(vocab + 1, dtype=torch.int64, device=dev)
    term_ptr[1:] = torch.cumsum(nd, 0)
    return dict(term_ptr=term_ptr, post_doc=(post["post_doc"][keep] - lo).to(torch.int32),
                post_tf=post["post_tf"][keep], doc_len=post["doc_len"][lo:hi].contiguous())


def test_config1_two_shards_replay_equals_unsharded(config1):
    """The sharded path at BASELINE size: the 1M corpus as TWO chunk shards in one process (500k
    rows + the same documents' postings each, global idf / avgdl) through anr_hybrid_search_keys
    -- the bf16 GEMM pass and the pruned BM25 scan, which the 2 048-row golden never reaches --,
    their keys laid out as the NCCL all-gather leaves them, then anr_sharded_fuse: bit-identical
    to the unsharded anr_hybrid_search (query_rag_retrieval.py:206-212, :308-315, :357-362)."""
    import torch
    c = config1
    dev = c["emb"].device
    c["dense"].set_shadow(True)
    whole = _hybrid(c, c["queries"], c["terms"])
    ctx = engine.context()
    q_dev = torch.from_numpy(c["queries"]).to(dev)
    terms, offsets = engine.Bm25Index.pack_queries([list(map(int, t)) for t in c["terms"]])
    t_dev, o_dev = torch.from_numpy(terms).to(dev), torch.from_numpy(offsets).to(dev)
    gathered = torch.zeros((2, 2, B, K), dtype=torch.int64, device=dev)
    bounds = [0, N // 2, N]
    keep_alive = []
    for s in range(2):
        lo, hi = bounds[s], bounds[s + 1]
        d_shard = engine.DenseIndex(c["emb"][lo:hi], borrow=True)
        d_shard.set_shadow(True)
        p = _shard_postings(torch, c["post"], VOCAB, lo, hi)
        b_shard = engine.Bm25Index(p["term_ptr"], p["post_doc"], p["post_tf"], p["doc_len"], c["idf"],
                                   1.7, 0.83, c["avgdl"], n_terms=VOCAB, n_docs=hi - lo)
        keep_alive += [d_shard, b_shard, p]
        native.call("anr_hybrid_search_keys", ctx.handle, d_shard.handle, b_shard.handle,
                    q_dev.data_ptr(), t_dev.data_ptr(), o_dev.data_ptr(), B, K, None, None, lo, lo,
                    gathered[s].data_ptr(), engine.torch_stream_ptr())
    ids = torch.empty((B, K), dtype=torch.int32, device=dev)
    scores = torch.empty((B, K), dtype=torch.float64, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    native.call("anr_sharded_fuse", ctx.handle, gathered.data_ptr(), 2, B, K, W_DENSE, W_BM25, WRRF_K,
                K, ids.data_ptr(), scores.data_ptr(), counts.data_ptr(), engine.torch_stream_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy(), whole["counts"])
    assert np.array_equal(ids.cpu().numpy(), whole["ids"])
    assert np.array_equal(scores.cpu().numpy(), whole["scores"])


def test_config3_bm25_10M_docs_batch256_properties():
    """BASELINE.json configs[3] at full size: BM25-only over a 10M-document CSR index (V = 500k,
    Zipf 1.1), 8-term queries, batch 256, top-10 -- the candidate-driven path (anr_bm25_ms.cu).
    Sampled queries against a float64 scatter of the same postings on the device (rank_bm25
    get_scores' formula, src/search_engine.py:219); the unpruned tiled scan of the same batch
    returns the same documents."""
    import os
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * (1 << 30):
        pytest.skip("needs ~60 GB of device memory")
    n, vocab, b = 10_000_000, 500_000, 256
    dev = torch.device("cuda", 0)
    post = synth.zipf_postings_torch(n, vocab, 1.1, 2024, dev)
    idf = synth.idf_from_counts(n, post["nd"].cpu().numpy(), 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / n
    index = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf,
                             1.7, 0.83, avgdl, n_terms=vocab, n_docs=n)
    tq = synth.zipf_queries(b, 8, vocab, 1.1, seed=2025)
    queries = [list(map(int, t)) for t in tq]
    scores, docs, counts = index.search(queries, K)
    assert (counts == K).all()
    tp, pdoc, ptf, dl = (post[k] for k in ("term_ptr", "post_doc", "post_tf", "doc_len"))
    for q in range(0, b, 37):
        acc = torch.zeros(n, dtype=torch.float64, device=dev)
        for term in tq[q]:
            lo, hi = int(tp[term]), int(tp[term + 1])
            dd = pdoc[lo:hi].long()
            tf = ptf[lo:hi].double()
            acc[dd] += idf[term] * (tf * 2.7 / (tf + 1.7 * (1 - 0.83 + 0.83 * dl[dd].double() / avgdl)))
        top = torch.topk(acc, K)
        got = torch.from_numpy(docs[q].astype(np.int64)).to(dev)
        assert torch.allclose(acc[got], top.values, rtol=1e-5, atol=1e-6), f"bm25 ranking q{q}"
        np.testing.assert_allclose(scores[q].astype(np.float64), acc[got].cpu().numpy(), rtol=1e-5,
                                   atol=1e-6)
        assert len(set(docs[q].tolist())) == K and np.all(np.diff(scores[q]) <= 0)
    os.environ["ANR_DISABLE_BM25_PRUNE"] = "1"
    try:
        s_all, d_all, _ = index.search(queries[:32], K)
    finally:
        del os.environ["ANR_DISABLE_BM25_PRUNE"]
    assert np.array_equal(d_all, docs[:32])
    np.testing.assert_allclose(s_all, scores[:32], rtol=2e-6, atol=1e-7)
