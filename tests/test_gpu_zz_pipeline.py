"""GPU: graph.HybridPipeline -- host buffers in and out with two (or three) batches in flight --
returns, in submission order, exactly what the synchronous C-ABI call returns."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
engine = importlib.import_module("a-nice-rag_b200.engine")
graph = importlib.import_module("a-nice-rag_b200.graph")
synth = importlib.import_module("a-nice-rag_b200.synth")

N, D, VOCAB, K = 100_000, 256, 20_000, 10
W_DENSE, W_BM25, WRRF_K = 5.0, 1.0, 40.0


@pytest.mark.parametrize("b,shadow,depth", [(1, True, 2), (64, True, 2), (40, False, 3)])
def test_pipeline_equals_synchronous_calls(b, shadow, depth):
    import torch
    dev = torch.device("cuda", 0)
    emb = synth.unit_vectors_torch(N, D, 31, dev)
    post = synth.zipf_postings_torch(N, VOCAB, 1.1, 32, dev)
    idf = synth.idf_from_counts(N, post["nd"].cpu().numpy(), 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / N
    dense = engine.DenseIndex(emb, borrow=True)
    dense.set_shadow(shadow)
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                            idf, 1.7, 0.83, avgdl, n_terms=VOCAB, n_docs=N)
    pipe = graph.HybridPipeline(dense, bm25, b, 8 * b, K, K, W_DENSE, W_BM25, WRRF_K, K, depth=depth)
    offsets = np.arange(0, 8 * b + 1, 8, dtype=np.int32)
    batches = [(synth.unit_vectors(b, D, seed=300 + i),
                synth.zipf_queries(b, 8, VOCAB, 1.1, seed=400 + i)) for i in range(7)]
    want = [engine.hybrid_search(dense, bm25, q, [list(map(int, t)) for t in terms], K, K, W_DENSE,
                                 W_BM25, WRRF_K, K) for q, terms in batches]
    got = []
    for i, (q, terms) in enumerate(batches):
        if i >= depth:
            got.append(tuple(a.copy() for a in pipe.collect()))
        pipe.submit(q, terms, offsets)
    while len(got) < len(batches):
        got.append(tuple(a.copy() for a in pipe.collect()))
    with pytest.raises(RuntimeError):
        pipe.collect()
    for i, ((ids, scores, counts), w) in enumerate(zip(got, want)):
        np.testing.assert_array_equal(counts, w["counts"], err_msg=f"batch {i}")
        np.testing.assert_array_equal(ids, w["ids"], err_msg=f"batch {i}")
        np.testing.assert_array_equal(scores, w["scores"], err_msg=f"batch {i}")
