"""GPU: a hybrid step captured as a CUDA graph (a-nice-rag_b200/graph.py) returns, replay after
replay and for inputs loaded AFTER the capture, exactly what the eager C-ABI call returns --
every dense path (fp32 CUDA-core scan, 32-query tcgen05 scan, tiled GEMM with tf32 operands and
with the bf16 shadow) beside the plain and the pruned BM25 scan, with and without filters."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
engine = importlib.import_module("a-nice-rag_b200.engine")
graph = importlib.import_module("a-nice-rag_b200.graph")
native = importlib.import_module("a-nice-rag_b200.native")
synth = importlib.import_module("a-nice-rag_b200.synth")

N, D, VOCAB, K = 100_000, 256, 20_000, 10
W_DENSE, W_BM25, WRRF_K = 5.0, 1.0, 40.0


@pytest.fixture(scope="module")
def corpus():
    import torch
    dev = torch.device("cuda", 0)
    emb = synth.unit_vectors_torch(N, D, 21, dev)
    post = synth.zipf_postings_torch(N, VOCAB, 1.1, 22, dev)
    idf = synth.idf_from_counts(N, post["nd"].cpu().numpy(), 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / N
    dense = engine.DenseIndex(emb, borrow=True)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                            idf, 1.7, 0.83, avgdl, n_terms=VOCAB, n_docs=N)
    return dict(emb=emb, dense=dense, bm25=bm25)


def _batch(b, seed):
    q = synth.unit_vectors(b, D, seed=seed)
    terms = synth.zipf_queries(b, 8, VOCAB, 1.1, seed=seed + 1)
    terms[0, 3] = -1                  # an out-of-vocabulary token
    return q, terms


@pytest.mark.parametrize("b,shadow,filtered", [(1, False, False), (1, True, False), (4, False, True),
                                               (20, False, False), (40, False, False),
                                               (40, True, True), (64, True, False)])
def test_graph_replay_equals_eager(corpus, b, shadow, filtered):
    import torch
    dense, bm25 = corpus["dense"], corpus["bm25"]
    dense.set_shadow(shadow)
    mask = None
    if filtered:
        mask = engine.pack_mask(np.random.default_rng(5).random(N) < 0.6)
    g = graph.HybridGraph(dense, bm25, b, 8 * b, K, K, W_DENSE, W_BM25, WRRF_K, K, row_mask=mask,
                          doc_mask=mask, want_lists=True)
    for seed in (100, 200, 100):      # new inputs after the capture, then the first ones again
        q, terms = _batch(b, seed)
        want = engine.hybrid_search(dense, bm25, q, [list(map(int, t)) for t in terms], K, K,
                                    W_DENSE, W_BM25, WRRF_K, K, row_mask=mask, doc_mask=mask,
                                    want_lists=True)
        offsets = np.arange(0, 8 * b + 1, 8, dtype=np.int32)
        ids, scores, counts = g.search(q, terms.astype(np.int32), offsets)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(counts.cpu().numpy(), want["counts"])
        np.testing.assert_array_equal(ids.cpu().numpy(), want["ids"])
        np.testing.assert_array_equal(scores.cpu().numpy(), want["scores"])
        for name, t in g.lists.items():
            np.testing.assert_array_equal(t.cpu().numpy(), want[name], err_msg=name)


def test_graph_rejects_wrong_shapes(corpus):
    g = graph.HybridGraph(corpus["dense"], corpus["bm25"], 2, 16, K, K, W_DENSE, W_BM25, WRRF_K, K)
    q, terms = _batch(2, 7)
    with pytest.raises(ValueError):
        g.load(q[:1], terms[:1].astype(np.int32), np.array([0, 8], dtype=np.int32))
    with pytest.raises(ValueError):
        g.load(q, np.zeros(17, dtype=np.int32), np.array([0, 8, 17], dtype=np.int32))


def test_consecutive_calls_on_different_streams_do_not_share_scratch_concurrently():
    """One context, two streams: an asynchronous device-pointer call on torch's stream followed at
    once by a host-pointer call (stream NULL = the context's own non-blocking stream).  Both use
    the context's scratch arena, so the library orders the second behind the first (found by the
    round-2 bench: a query of the parity batch lost a hit when the two overlapped)."""
    import torch
    n, d, vocab, b, k = 300_000, 256, 20_000, 64, 10
    dev = torch.device("cuda", 0)
    emb = synth.unit_vectors_torch(n, d, 71, dev)
    post = synth.zipf_postings_torch(n, vocab, 1.1, 72, dev)
    idf = synth.idf_from_counts(n, post["nd"].cpu().numpy(), 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / n
    dense = engine.DenseIndex(emb, borrow=True)
    dense.set_shadow(True)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf,
                            1.7, 0.83, avgdl, n_terms=vocab, n_docs=n)
    qa, qb = synth.unit_vectors(b, d, seed=73), synth.unit_vectors(b, d, seed=74)
    ta = [list(map(int, t)) for t in synth.zipf_queries(b, 8, vocab, 1.1, seed=75)]
    tb = [list(map(int, t)) for t in synth.zipf_queries(b, 8, vocab, 1.1, seed=76)]
    want_a = engine.hybrid_search(dense, bm25, qa, ta, k, k, 5.0, 1.0, 40.0, k, want_lists=True)
    want_b = engine.hybrid_search(dense, bm25, qb, tb, k, k, 5.0, 1.0, 40.0, k, want_lists=True)
    terms, offsets = engine.Bm25Index.pack_queries(ta)
    q_dev = torch.from_numpy(qa).to(dev)
    t_dev, o_dev = torch.from_numpy(terms).to(dev), torch.from_numpy(offsets).to(dev)
    ids = torch.empty((b, k), dtype=torch.int32, device=dev)
    scores = torch.empty((b, k), dtype=torch.float64, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    ctx = engine.context()
    torch.cuda.synchronize()
    for _ in range(10):
        native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle, q_dev.data_ptr(),
                    t_dev.data_ptr(), o_dev.data_ptr(), b, k, k, None, None, None, 0, 5.0, 1.0, 40.0, k,
                    ids.data_ptr(), scores.data_ptr(), counts.data_ptr(), None, None, None, None,
                    engine.torch_stream_ptr())
        got_b = engine.hybrid_search(dense, bm25, qb, tb, k, k, 5.0, 1.0, 40.0, k, want_lists=True)
        torch.cuda.synchronize()
        for key in ("ids", "scores", "counts", "dense_rows", "bm25_ids"):
            assert np.array_equal(got_b[key], want_b[key]), key
        assert np.array_equal(ids.cpu().numpy(), want_a["ids"])
        assert np.array_equal(scores.cpu().numpy(), want_a["scores"])
