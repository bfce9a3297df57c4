"""BASELINE.json configs[2] and configs[3] on one B200 (records for profiles/, not bench lines).

  config3: 10M x 1024 fp32, batch 1024, dense-only top-100 through anr_dense_search
           (tcgen05 tf32 scan, 64 queries per pass, exact fp32 rescoring); parity of a sample of
           queries against torch fp32 matmul (allow_tf32 off) + topk on the same device tensor.
  config4: BM25-only over a 10M-document CSR index (V = 500k, Zipf 1.1), 8-term queries,
           batch 256, top-10 through anr_bm25_search; algorithmic bytes = 8 B x sum of df.
Usage: python profiles/config_bench.py [config3|config4] [n_rows]
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("a-nice-rag_b200")
engine, native, synth = pkg.engine, pkg.native, importlib.import_module("a-nice-rag_b200.synth")

which = sys.argv[1] if len(sys.argv) > 1 else "config3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
dev = torch.device("cuda", 0)
ctx = engine.context(0)
PEAK = 6550.7
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


if which == "config3":
    d, b, k = 1024, 1024, 100
    emb = synth.unit_vectors_torch(n, d, 1234, dev)
    index = engine.DenseIndex(emb, borrow=True)
    q = torch.from_numpy(synth.unit_vectors(b, d, seed=4321)).to(dev)
    scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    rows = torch.empty((b, k), dtype=torch.int32, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    stream = engine.torch_stream_ptr()

    def run():
        native.call("anr_dense_search", ctx.handle, index.handle, q.data_ptr(), b, k, None, 0,
                    scores.data_ptr(), rows.data_ptr(), counts.data_ptr(), stream)

    ms = timed(run, 3, warm=1)
    # parity sample against torch fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    checked = 0
    for qi in range(0, b, 128):
        ref = torch.mv(emb, q[qi])
        top = torch.topk(ref, k)
        got_rows = rows[qi].long()
        same = bool((got_rows == top.indices).all())
        if not same:   # ties / last-ulp differences: every returned row must score like the ref row
            diff = (ref[got_rows] - top.values).abs().max().item()
            assert diff <= 1e-5 * top.values.abs().max().item() + 1e-6, (qi, diff)
        assert torch.allclose(scores[qi], ref[got_rows], rtol=1e-5, atol=1e-6)
        checked += 1
    passes = (b + 63) // 64
    print(json.dumps({
        "config": f"{n} x {d} fp32, batch {b}, dense-only top-{k} (tcgen05 tf32 + fp32 rescoring)",
        "ms_per_batch": ms, "queries_per_s": b / ms * 1e3, "corpus_passes": passes,
        "hbm_gbs_algorithmic": passes * n * d * 4 / ms / 1e6, "frac_of_measured_hbm": passes * n * d * 4 / ms / 1e6 / PEAK,
        "tf32_tflops": 2.0 * b * n * d / ms / 1e9, "parity_checked_queries": checked}))
else:
    vocab, b, t, k = 500_000, 256, 8, 10
    post = synth.zipf_postings_torch(n, vocab, 1.1, 2024, dev)
    nd = post["nd"].cpu().numpy()
    idf = synth.idf_from_counts(n, nd, 0.05)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / n
    index = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                             idf, 1.7, 0.83, avgdl, n_terms=vocab, n_docs=n)
    tq = synth.zipf_queries(b, t, vocab, 1.1, seed=2025)
    df_sum = float(nd[tq].sum())
    terms = torch.from_numpy(tq.reshape(-1).copy()).to(dev)
    offs = torch.arange(0, (b + 1) * t, t, dtype=torch.int32, device=dev)
    scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    docs = torch.empty((b, k), dtype=torch.int32, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    stream = engine.torch_stream_ptr()

    def run():
        native.call("anr_bm25_search", ctx.handle, index.handle, terms.data_ptr(), offs.data_ptr(),
                    b, k, None, None, 0, scores.data_ptr(), docs.data_ptr(), counts.data_ptr(),
                    stream)

    ms = timed(run, 5)
    # parity of a few queries against a torch float64 scatter of the same postings
    tp = post["term_ptr"]
    checked = 0
    for qi in range(0, b, 64):
        acc = torch.zeros(n, dtype=torch.float64, device=dev)
        for term in tq[qi]:
            lo, hi = int(tp[term]), int(tp[term + 1])
            dd = post["post_doc"][lo:hi].long()
            tf = post["post_tf"][lo:hi].double()
            dl = post["doc_len"][dd].double()
            acc[dd] += idf[term] * (tf * 2.7 / (tf + 1.7 * (1 - 0.83 + 0.83 * dl / avgdl)))
        top = torch.topk(acc, k)
        got = docs[qi].long()
        diff = (acc[got] - top.values).abs().max().item()
        assert diff <= 1e-5 * top.values.abs().max().item() + 1e-6, (qi, diff)
        assert torch.allclose(scores[qi].double(), acc[got], rtol=1e-5, atol=1e-6)
        checked += 1
    print(json.dumps({
        "config": f"BM25-only, {n} docs, V={vocab}, Zipf 1.1, {t}-term queries, batch {b}, top-{k}",
        "postings": index.n_postings, "ms_per_batch": ms, "queries_per_s": b / ms * 1e3,
        "sum_df_per_query": df_sum / b, "algorithmic_bytes_per_batch": 8 * df_sum,
        "hbm_gbs_algorithmic": 8 * df_sum / ms / 1e6, "frac_of_measured_hbm": 8 * df_sum / ms / 1e6 / PEAK,
        "parity_checked_queries": checked}))
