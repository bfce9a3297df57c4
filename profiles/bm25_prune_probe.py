"""Times BM25 batch search (anr_bm25_search) with the unpruned and the pruned scan on the
bench corpus (Zipf 1.1, 8-term queries) and checks that both agree.
Usage: python profiles/bm25_pruned_probe.py <n_docs> <vocab> <batch> [iters]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
b = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
t, k = 8, 10
dev = torch.device("cuda", 0)
ctx = engine.context(0)
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
stream = engine.torch_stream_ptr()
native.call("anr_ctx_profile_enable", ctx.handle, 1)
out = {}
for mode in ("pruned", "unpruned"):
    if mode == "unpruned":
        os.environ["ANR_DISABLE_BM25_PRUNE"] = "1"
    scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    docs = torch.empty((b, k), dtype=torch.int32, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)

    def run():
        native.call("anr_bm25_search", ctx.handle, index.handle, terms.data_ptr(), offs.data_ptr(),
                    b, k, None, None, 0, scores.data_ptr(), docs.data_ptr(), counts.data_ptr(),
                    stream)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    native.call("anr_ctx_profile_read", ctx.handle, 1, None, None)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        run()
    e.record()
    torch.cuda.synchronize()
    ms, cnt = C.c_double(), C.c_int64()
    native.call("anr_ctx_profile_read", ctx.handle, 1, C.byref(ms), C.byref(cnt))
    out[mode] = dict(call_ms=s.elapsed_time(e) / iters, kernel_ms=ms.value / max(cnt.value, 1),
                     scores=scores.cpu().numpy(), docs=docs.cpu().numpy())
g, p = out["pruned"], out["unpruned"]
same_docs = float((g["docs"] == p["docs"]).mean())
max_rel = float(np.max(np.abs(g["scores"] - p["scores"]) / np.maximum(np.abs(p["scores"]), 1e-9)))
print(json.dumps({
    "n_docs": n, "vocab": vocab, "batch": b, "postings": int(post["term_ptr"][-1]),
    "sum_df_unpruned": df_sum / b, "algorithmic_bytes_per_batch": 8 * df_sum,
    "pruned_kernel_ms": g["kernel_ms"], "pruned_call_ms": g["call_ms"],
    "unpruned_kernel_ms": p["kernel_ms"], "unpruned_call_ms": p["call_ms"],
    "pruned_gbs_algorithmic": 8 * df_sum / g["kernel_ms"] / 1e6,
    "unpruned_gbs_algorithmic": 8 * df_sum / p["kernel_ms"] / 1e6,
    "same_docs_fraction": same_docs, "max_rel_score_diff": max_rel}), flush=True)
