"""CPU, world_size 2 over gloo: the host side of the sharded path -- shard ranges, the all-reduce
that gives every shard GLOBAL BM25 statistics, and the key packing contract of the merge
(score-ordered 64-bit keys with global ids: merging per-shard top-k lists by key reproduces the
unsharded ranking)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def f32_to_ord(x: np.ndarray) -> np.ndarray:
    u = (x.astype(np.float32) + np.float32(0.0)).view(np.uint32).astype(np.uint64)
    neg = (u >> np.uint64(31)) != 0
    return np.where(neg, (~u) & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))


def make_keys(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """The library's key format (anr_common.cuh): ordered(score) << 32 | (0xFFFFFFFF - id)."""
    return (f32_to_ord(scores) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - ids.astype(np.uint64))


def _worker(rank: int, world: int, port: int, tmp: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("a-nice-rag_b200.sharded")
    synth = importlib.import_module("a-nice-rag_b200.synth")
    from oracle import csr, retrieval

    n, vocab, d, k = 3001, 300, 32, 10
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, 1.1, seed=5, len_lo=10, len_hi=40)
    emb = synth.unit_vectors(n, d, seed=6)
    lo, hi = sharded.shard_range(n, rank, world)

    # ---- global BM25 statistics from local document frequencies ----
    local = csr.from_token_ids(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[doc_ptr[lo]:doc_ptr[hi]],
                               vocab, 1.7, 0.83, 0.05)
    nd_local = torch.from_numpy(np.diff(local.term_ptr))
    nd, n_total, avgdl = sharded.global_bm25_stats(nd_local, int(local.doc_len.sum()), hi - lo)
    full = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    assert n_total == n and np.array_equal(nd.numpy(), np.diff(full.term_ptr))
    assert abs(avgdl - full.avgdl) < 1e-12
    idf = synth.idf_from_counts(n_total, nd.numpy(), 0.05)
    np.testing.assert_allclose(idf, full.idf, rtol=1e-12, atol=1e-15)

    # ---- shard scores with global stats == unsharded scores on the shard's documents ----
    local.idf, local.avgdl = idf, avgdl
    terms = [0, 3, 3, 17, 120]
    np.testing.assert_allclose(csr.scores(local, terms), csr.scores(full, terms)[lo:hi],
                               rtol=1e-12, atol=0)

    # ---- key exchange: all-gather local top-k keys, merge by key == unsharded top-k ----
    q = synth.unit_vectors(1, d, seed=7)[0]
    scores_local = emb[lo:hi] @ q
    top = retrieval.topk_desc(scores_local, k)
    keys = make_keys(scores_local[top], top + lo)
    gathered = [torch.zeros(k, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(keys.view(np.int64)))
    merged = np.sort(np.concatenate([g.numpy().view(np.uint64) for g in gathered]))[::-1][:k]
    ids = (np.uint64(0xFFFFFFFF) - (merged & np.uint64(0xFFFFFFFF))).astype(np.int64)
    want = retrieval.topk_desc(emb @ q, k)
    assert ids.tolist() == want.tolist()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").close()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
