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
