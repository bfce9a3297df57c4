"""Python model of the candidate-driven BM25 top-k (a-nice-rag_b200/csrc/anr_bm25_ms.cu).
TEST INFRASTRUCTURE ONLY.

It restates, step by step, the decisions the CUDA path takes for ONE query -- sample selection,
theta from the sample, the required set, the quick and progressive bound tests, the
ownership rule (shortest required list first) that keeps a document from being listed twice, the conditions that flag a query
for the exhaustive scan -- with fp32 arithmetic, so that ``tests/test_oracle.py`` can prove on the
CPU that the pruning is SAFE: whenever the model does not flag a query, its survivors contain the
exact top-k of the exhaustive BM25 scores (``oracle/csr.py``, itself pinned to BM25Okapi).
What it cannot prove is that the CUDA code follows it; that is what the GPU parity tests are for.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

MAX_TERMS, SAMPLE, SURVIVORS, THREADS = 48, 4096, 4096, 256
f32 = np.float32


def posting_weights(ix) -> np.ndarray:
    """fp32 posting weights as the library builds them (float64 formula, rounded once)."""
    tf = ix.post_tf.astype(np.float64)
    dl = ix.doc_len[ix.post_doc].astype(np.float64)
    return (tf * (ix.k1 + 1) / (tf + ix.k1 * (1 - ix.b + ix.b * dl / ix.avgdl))).astype(np.float32)


def topk(ix, post_w: np.ndarray, term_ids: Sequence[int], k: int,
         allowed: Optional[np.ndarray] = None, head_df: Optional[int] = None):
    """-> (flagged, [(score, doc)] survivors sorted best first, stats dict)."""
    n = len(term_ids)
    if n == 0:
        return False, [], {}
    if n > MAX_TERMS:
        return True, [], {}
    idf32 = ix.idf.astype(np.float32)
    lo = np.zeros(n, np.int64); ln = np.zeros(n, np.int64)
    idf = np.zeros(n, np.float32); ub = np.zeros(n, np.float32)
    for j, t in enumerate(term_ids):
        if 0 <= t < len(ix.idf) and idf32[t] > 0:
            a, b = int(ix.term_ptr[t]), int(ix.term_ptr[t + 1])
            if b > a:
                lo[j], ln[j], idf[j] = a, b - a, idf32[t]
                ub[j] = f32(idf32[t] * post_w[a:b].max())
    tot = f32(0)
    for j in range(n):
        tot = f32(tot + ub[j])
    slack = f32(2e-5) * tot
    # ---- evaluation order of stage 2: largest bound first; esuf[r] = bounds of ranks >= r
    ord_ = sorted(range(n), key=lambda j: (-ub[j], j))
    rank = {j: r for r, j in enumerate(ord_)}
    esuf = np.zeros(n + 1, np.float32)
    for r in range(n - 1, -1, -1):          # (the kernel sums in another order: bounds carry slack)
        esuf[r] = f32(esuf[r + 1] + ub[ord_[r]])
    # ---- sample: shortest lists first, SAMPLE postings in all
    order = sorted((j for j in range(n) if ln[j] > 0), key=lambda j: (ln[j], j))
    s1 = np.zeros(n, np.int64)
    room = SAMPLE
    for j in order:
        s1[j] = min(ln[j], room)
        room -= s1[j]

    def weight(j: int, doc: int) -> Tuple[float, int]:
        docs = ix.post_doc[lo[j]:lo[j] + ln[j]]
        p = int(np.searchsorted(docs, doc))
        if p < ln[j] and docs[p] == doc:
            return post_w[lo[j] + p], p
        return f32(0), -1

    def exact(j_self: int, c_self, doc: int):
        full = f32(0)
        for jj in range(n):                 # query order, fp32 fma (duplicates repeat)
            if ln[jj] > 0:
                w = c_self if jj == j_self else weight(jj, doc)[0]
                if w > 0:
                    full = f32(np.float64(idf[jj]) * np.float64(w) + np.float64(full))
        return full

    def full_score(j_self: int, p_self: int, stage: int, theta) -> Optional[np.float32]:
        doc = int(ix.post_doc[lo[j_self] + p_self])
        if allowed is not None and not allowed[doc]:
            return None
        c_self = post_w[lo[j_self] + p_self]
        if stage == 1:
            for jj in range(j_self):        # sampled through an earlier list: that list's candidate
                if s1[jj] > 0:
                    p = weight(jj, doc)[1]
                    if 0 <= p < s1[jj]:
                        return None
            return exact(j_self, c_self, doc)
        mine = f32(idf[j_self] * c_self)

        def before(jj: int) -> bool:        # required lists, shortest first: who owns a document
            return s2[jj] > 0 and (s2[jj] < s2[j_self] or (s2[jj] == s2[j_self] and jj < j_self))
        sub = f32(0)
        for jj in range(n):
            if before(jj):
                sub = f32(sub + ub[jj])
        remaining = f32(f32(tot - ub[j_self]) - sub)
        if f32(f32(mine + remaining) + slack) < theta:
            return None
        partial = mine
        for r in range(n):                  # the terms that can add, largest bound first
            jj = ord_[r]
            if ln[jj] == 0:
                break
            if jj == j_self or before(jj):
                continue
            remaining = f32(remaining - ub[jj])
            w = weight(jj, doc)[0]
            if w > 0:
                partial = f32(np.float64(idf[jj]) * np.float64(w) + np.float64(partial))
            if f32(f32(partial + max(remaining, f32(0))) + slack) < theta:
                return None
        for jj in range(n):                 # then: does a list ordered before mine own the document?
            if jj != j_self and before(jj) and weight(jj, doc)[1] >= 0:
                return None
        return exact(j_self, c_self, doc)

    # ---- stage 1 + theta (k-th best full score of the sample)
    s2 = np.zeros(n, np.int64)
    keys: List[Tuple[float, int]] = []
    for j in range(n):
        for p in range(int(s1[j])):
            s = full_score(j, p, 1, f32(0))
            keys.append((float(s), -int(ix.post_doc[lo[j] + p])) if s is not None and s > 0 else (0.0, 0))
    # (the kernel: k-th largest of the 256 per-thread bests as a first bound, then the exact k-th
    #  best among the keys at or above it -- the k-th best of the sample)
    ranked = sorted(keys, reverse=True)
    theta = f32(ranked[k - 1][0]) if len(ranked) >= k and ranked[k - 1][0] > 0 else f32(0)
    # ---- required set: set aside, longest list first, while the bounds stay below theta
    cum = f32(0)
    aside = set()
    for j in sorted((j for j in range(n) if ln[j] > 0), key=lambda j: (-ln[j], j)):
        cum = f32(cum + ub[j])
        if f32(cum + slack) < theta:
            aside.add(j)
        else:
            break
    for j in range(n):
        s2[j] = ln[j] if (ln[j] > 0 and j not in aside) else 0
    # ---- stage 2
    surv = []
    for j in range(n):
        for p in range(int(s2[j])):
            s = full_score(j, p, 2, theta)
            if s is not None and s >= theta and s > 0:
                surv.append((float(s), int(ix.post_doc[lo[j] + p])))
    flagged = len(surv) < k or len(surv) > SURVIVORS
    surv.sort(key=lambda sd: (-sd[0], sd[1]))
    return flagged, surv, dict(theta=float(theta), streamed=int(s2.sum()), sampled=int(s1.sum()),
                               aside=len(aside), survivors=len(surv))
