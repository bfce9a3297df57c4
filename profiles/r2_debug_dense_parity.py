#!/usr/bin/env python
"""Debug: which dense results differ from torch's exact fp32 top-10 on the 1M bench corpus, per
code path (dense-only bf16 / tf32 / hybrid with the capped ring), and how stable that is."""
import importlib, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("a-nice-rag_b200")
engine, native = pkg.engine, pkg.native
synth = importlib.import_module("a-nice-rag_b200.synth")
N, D, B, K = 1_000_000, 1024, 64, 10
dev = torch.device("cuda", 0)
emb = synth.unit_vectors_torch(N, D, 1234, dev)
q_host = synth.unit_vectors(B, D, seed=4321)
q = torch.from_numpy(q_host).to(dev)
torch.backends.cuda.matmul.allow_tf32 = False
ref = (emb @ q.T).T.contiguous()              # [B, N] fp32
top = torch.topk(ref, K, dim=1)
dense = engine.DenseIndex(emb, borrow=True)

def check(name, rows, scores):
    bad = []
    for b in range(B):
        want = top.indices[b].cpu().numpy()
        if not np.array_equal(np.sort(rows[b]), np.sort(want)):
            miss = sorted(set(want.tolist()) - set(rows[b].tolist()))
            extra = sorted(set(rows[b].tolist()) - set(want.tolist()))
            info = []
            for r in miss:
                e16 = emb[r].to(torch.bfloat16).float(); q16 = q[b].to(torch.bfloat16).float()
                info.append(dict(row=r, tile=r // 128, in_tile=r % 128, exact=float(ref[b, r]),
                                 bf16=float((e16 * q16).sum()), rank=int((ref[b] > ref[b, r]).sum())))
            bad.append(dict(q=b, missing=info, extra=extra, kth=float(top.values[b, -1])))
    print(name, "mismatching queries:", len(bad), json.dumps(bad)[:1500], flush=True)
    return len(bad)

for shadow in (True, False):
    dense.set_shadow(shadow)
    for rep in range(3):
        s, r, c = dense.search(q_host, K)
        check(f"dense-only shadow={shadow} rep{rep}", r, s)
# hybrid path (ring cap) with a BM25 index
post = synth.zipf_postings_torch(N, 50_000, 1.1, 2024, dev)
idf = synth.idf_from_counts(N, post["nd"].cpu().numpy(), 0.05)
avgdl = float(post["doc_len"].to(torch.int64).sum()) / N
bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"], idf, 1.7, 0.83,
                        avgdl, n_terms=50_000, n_docs=N)
tq = synth.zipf_queries(B, 8, 50_000, 1.1, seed=2025)
dense.set_shadow(True)
for rep in range(3):
    got = engine.hybrid_search(dense, bm25, q_host, [list(map(int, t)) for t in tq], K, K, 5.0, 1.0, 40.0, K,
                               want_lists=True)
    check(f"hybrid shadow rep{rep}", got["dense_rows"], got["dense_scores"])
# batch sizes around the tile widths
for b in (1, 8, 33, 63):
    s, r, c = dense.search(q_host[:b], K)
    bad = sum(not np.array_equal(np.sort(r[i]), np.sort(top.indices[i].cpu().numpy())) for i in range(b))
    print("dense-only shadow batch", b, "bad", bad, flush=True)
