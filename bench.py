#!/usr/bin/env python
"""Benchmark of the retrieval hot path: hybrid queries/s (dense top-10 + BM25 top-10 + weighted
RRF -> top-10) on BASELINE.json configs[1] -- 1M chunks x 1024-d fp32 + a Zipf(1.1) BM25 corpus of
1M documents (V=50k, 8-term queries) -- in batches of 64 queries, plus batch-1 latency, plus one
bounded leg per other BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs a,b,..]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A step = one batch of B hybrid queries.
  value      queries/s, inputs resident in HBM, CUDA events on the launching stream, max over
             ranks: the MEDIAN of `--blocks` back-to-back blocks of exactly K steps (`blocks` holds
             min / max / every block)
  e2e        the same through the C ABI with pinned HOST buffers (H2D of the queries and D2H of
             the fused results inside the timed region)
  roofline   the kernel with the largest share of the step: algorithmic bytes per launch / mean
             launch time (events on the launching stream, live in this run) against
             MEASURED_PEAKS.json; `traffic` = ncu dram bytes per launch (profiles/traffic.json),
             `frac_traffic` = traffic / time / peak
  cpu_baseline  the CPU oracle port (numpy BLAS dot + CSR BM25 + Python RRF on a pre-stacked
             matrix) on the batch's queries over the full corpus, on this box's host cores; the
             SAME results are the parity check of every benchmark query (`parity_checked_queries`)
  legs       config3 (10M x 1024, batch 1024, top-100 dense), config4 (BM25-only, 10M docs,
             V=500k, batch 256), batch1_10M (batch-1 hybrid at 10M: the >= 70 % HBM target), weak
             (12.5M chunks per GPU -> 100M on 8 GPUs), each with its own parity count and clocks
N > 1: the SAME 1M-chunk corpus is sharded by chunk over the ranks (strong scaling, the headline
line); local top-k keys are exchanged with one NCCL all-gather and merged + fused on every rank;
`multi_gpu` times the three parts.  Rank 0 re-creates every shard for the CPU port, so the
all-gathered result of every benchmark query is checked at every N.
--impl reference: the CPU port alone (rank 0), timed on the host cores; this arm does not import
the product package, so none of the CUDA library is loaded.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hybrid queries/sec (1M x 1024-d + BM25, top-10)"
D = 1024
ZIPF_S = 1.1
K1, B_PARAM, EPS = 1.7, 0.83, 0.05
W_DENSE, W_BM25, WRRF_K = 5.0, 1.0, 40.0
WEIGHTS = {"voyage-3-large": W_DENSE, "BM25": W_BM25}
TOPK = 10
N_TERMS = 8
Q_SEED, T_SEED = 4321, 2025
EMB_SEED, POST_SEED = 1234, 2024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=1_000_000, help="total corpus rows / BM25 docs")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--vocab", type=int, default=50_000)
    ap.add_argument("--chunks-per-gpu", type=int, default=0,
                    help="headline corpus = this x world size (overrides --chunks)")
    ap.add_argument("--blocks", type=int, default=11, help="timed blocks of --steps steps each")
    ap.add_argument("--cpu-queries", type=int, default=64,
                    help="queries of the batch run through the CPU port (timed + parity)")
    ap.add_argument("--latency-iters", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-operands", default="bf16", choices=["bf16", "tf32"],
                    help="operands of the tensor-core nomination pass: a bf16 shadow copy of the "
                         "matrix (half the HBM bytes) or the fp32 words read as tf32; results are "
                         "identical (exact fp32 rescoring)")
    ap.add_argument("--legs", default="headline,big,weak",
                    help="comma list of: headline (always), big (10M legs, 1 GPU), weak "
                         "(12.5M chunks per GPU)")
    ap.add_argument("--big-chunks", type=int, default=10_000_000)
    ap.add_argument("--big-vocab", type=int, default=500_000)
    ap.add_argument("--weak-chunks-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--leg-steps", type=int, default=10, help="timed steps of each extra leg")
    ap.add_argument("--leg-budget-s", type=float, default=420.0,
                    help="wall-clock budget of the extra legs; past it the headline line is "
                         "printed without them")
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def load_synth():
    """synth.py by PATH: importing it as a-nice-rag_b200.synth would import the package, whose
    __init__ loads the CUDA library -- the reference arm must not do that."""
    spec = importlib.util.spec_from_file_location(
        "anr_bench_synth", os.path.join(ROOT, "a-nice-rag_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured")
    # B200_PROFILING.md fallback figures
    return dict(hbm=6650.0, bf16=1700.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while a timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self) -> int:
        return len(self.lines)

    def summary(self, since: int = 0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[since:]:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


# ---------------------------------------------------------------------------------------
# synthetic workload (device generation; identical bytes are copied to the host for the CPU arm)
# ---------------------------------------------------------------------------------------
def make_queries(synth, batch: int, vocab: int):
    q = synth.unit_vectors(batch, D, seed=Q_SEED)
    terms = synth.zipf_queries(batch, N_TERMS, vocab, ZIPF_S, seed=T_SEED)
    offsets = np.arange(0, (batch + 1) * N_TERMS, N_TERMS, dtype=np.int32)
    return q, terms, offsets


def make_shard(synth, device, lo: int, hi: int, shard_id: int, vocab: int, want_emb=True,
               want_post=True):
    emb = synth.unit_vectors_torch(hi - lo, D, EMB_SEED + shard_id, device) if want_emb else None
    post = (synth.zipf_postings_torch(hi - lo, vocab, ZIPF_S, POST_SEED + shard_id, device)
            if want_post else None)
    return emb, post


def workload_name(chunks: int, vocab: int) -> str:
    return (f"{chunks} chunks x {D}-d fp32 + BM25 {chunks} docs V={vocab} Zipf {ZIPF_S}, "
            f"{N_TERMS}-term queries, top-{TOPK} WRRF (w 5:1, k=40)")


def headline_config(args) -> dict:
    """The `config` object: identical in both arms (everything in it follows from the arguments)."""
    return {"workload": workload_name(args.chunks, args.vocab), "batch": args.batch,
            "sharding": f"chunks sharded over {args.gpus} GPU(s), queries replicated",
            "l2": "inputs larger than L2 (corpus >= 0.5 GB per GPU, re-read every step)"}


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port, timed (and used as the checker of the GPU results)
# ---------------------------------------------------------------------------------------
def set_cpu_threads() -> int:
    """All host cores for the BLAS pool, whatever OMP_NUM_THREADS the launcher exported (torchrun
    sets it to 1).  Returns the BLAS thread count actually in force."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        n = max((int(p.get("num_threads", 0)) for p in threadpool_info()
                 if p.get("user_api") == "blas"), default=0)
        if n > 0:
            return n
    except Exception:
        pass
    return want


def cpu_index(post_host, idf, avgdl):
    from oracle import csr
    return csr.CsrIndex(term_ptr=post_host["term_ptr"], post_doc=post_host["post_doc"],
                        post_tf=post_host["post_tf"],
                        doc_len=post_host["doc_len"].astype(np.int64), idf=idf, avgdl=avgdl,
                        k1=K1, b=B_PARAM)


def cpu_port_run(shards, queries, terms, which):
    """Runs the hybrid queries `which` on the CPU; returns (seconds, {q: result}).  shards =
    [(first row, emb, CsrIndex)] in row order (one entry = the unsharded corpus)."""
    from oracle import pipeline
    results = {}
    t0 = time.perf_counter()
    for q in which:
        tq = [int(t) for t in terms[q]]
        if len(shards) == 1:
            results[q] = pipeline.hybrid_query(queries[q], shards[0][1], shards[0][2], tq, TOPK, TOPK,
                                               WEIGHTS, WRRF_K, TOPK)
        else:
            results[q] = pipeline.hybrid_query_sharded(queries[q], shards, tq, TOPK, TOPK, WEIGHTS,
                                                       WRRF_K, TOPK)
    return time.perf_counter() - t0, results


def check_against_cpu(results, got, shards, queries):
    """GPU output vs the oracle for every query in `results`: the two ranked lists tie-aware (ids
    AND the scores the GPU reported, 1e-5 relative; the absolute 1e-6 only guards scores near 0),
    the fusion exactly given the lists.  Returns the number of queries checked."""
    from oracle import pipeline, retrieval
    for q, ref in results.items():
        d_all = ref.get("dense_all")
        if d_all is None and not np.array_equal(got["dense_rows"][q, :TOPK], ref["dense_ids"]):
            d_all = np.concatenate([retrieval.dense_scores(queries[q], emb) for _, emb, _ in shards])
        retrieval.assert_ranking_matches(
            got["dense_rows"][q, :TOPK], got["dense_scores"][q, :TOPK], ref["dense_ids"],
            ref["dense_scores"], all_scores=d_all, rtol=1e-5, atol=1e-6, what=f"bench dense q{q}")
        retrieval.assert_ranking_matches(
            got["bm25_ids"][q, :TOPK], got["bm25_scores"][q, :TOPK], ref["bm25_docs"],
            ref["bm25_all"][ref["bm25_docs"]], all_scores=ref["bm25_all"], rtol=1e-5, atol=1e-6,
            what=f"bench bm25 q{q}")
        c = int(got["counts"][q])
        pipeline.check_fused(got["ids"][q, :c], got["scores"][q, :c], got["dense_rows"][q, :TOPK],
                             got["bm25_ids"][q, :TOPK], (W_DENSE, W_BM25), WRRF_K, TOPK)
    return len(results)


def stock_path_config0(n_queries: int = 8):
    """Context figure for `cpu_baseline`: the reference's STOCK per-query path on BASELINE
    configs[0] (20k chunks x 1024-d, V=50k) restated literally -- np.stack of the object column on
    every call (src/search_engine.py:80), the rank_bm25 get_scores loop (:219), the Python RRF --
    next to the pre-stacked / CSR port the baseline line times.  (The unmodified reference itself,
    timed where it is mounted: profiles/r1_config0_reference_cpu.json, 7.5 queries/s.)"""
    from oracle import bm25_okapi, retrieval
    synth = load_synth()
    n, vocab = 20_000, 50_000
    emb = synth.unit_vectors(n, D, seed=EMB_SEED)
    column = np.empty(n, dtype=object)
    for i in range(n):
        column[i] = emb[i]
    doc_ptr, tokens = synth.zipf_corpus(n, vocab, ZIPF_S, seed=POST_SEED)
    okapi = bm25_okapi.BM25Okapi(synth.doc_token_lists(doc_ptr, tokens), k1=K1, b=B_PARAM,
                                 epsilon=EPS)
    queries = synth.unit_vectors(n_queries, D, seed=Q_SEED)
    tq = synth.zipf_queries(n_queries, N_TERMS, vocab, ZIPF_S, seed=T_SEED)
    t0 = time.perf_counter()
    for q in range(n_queries):
        stacked = np.stack(column)
        scores = np.dot(queries[q].reshape(1, -1), stacked.T).flatten()
        d_ids = retrieval.topk_desc(scores, TOPK)
        b_scores = okapi.get_scores(synth.token_strings(tq[q]))
        b_ids = retrieval.topk_desc(np.array(b_scores), TOPK)
        retrieval.weighted_rrf([([int(i) for i in d_ids], "voyage-3-large"),
                                ([int(i) for i in b_ids], "BM25")], WEIGHTS, WRRF_K)[:TOPK]
    dt = time.perf_counter() - t0
    return {"value": n_queries / dt, "unit": "queries/s",
            "what": f"stock per-query path on configs[0] ({n} chunks, V={vocab}): np.stack per "
                    f"call + literal BM25Okapi.get_scores + Python RRF, {n_queries} queries"}


# ---------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU port on this box's host cores; rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import torch
    synth = load_synth()
    cores = set_cpu_threads()
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    emb, post = make_shard(synth, device, 0, args.chunks, 0, args.vocab)   # torch only generates
    emb_host = emb.cpu().numpy()
    post_host = {k: v.cpu().numpy() for k, v in post.items()}
    del emb, post
    idf = synth.idf_from_counts(args.chunks, post_host["nd"], EPS)
    avgdl = float(post_host["doc_len"].astype(np.int64).sum() / args.chunks)
    queries, terms, _ = make_queries(synth, args.batch, args.vocab)
    shards = [(0, emb_host, cpu_index(post_host, idf, avgdl))]
    per_step = max(1, min(8, args.batch))
    for _ in range(min(args.warmup, 3)):
        cpu_port_run(shards, queries, terms, [0])
    total, done = 0.0, 0
    for s in range(args.steps):
        which = [(s * per_step + i) % args.batch for i in range(per_step)]
        dt, _ = cpu_port_run(shards, queries, terms, which)
        total += dt
        done += per_step
    qps = done / total
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} of the batch's {args.batch} queries per step over "
                                   "the full corpus; numpy BLAS dot + CSR BM25 + Python RRF on a "
                                   "pre-stacked matrix"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------
class Env:
    pass


class Workload:
    """One rank's shard of a corpus: device tensors + the two indices."""

    def __init__(self, env, n_total: int, vocab: int, shadow: bool):
        torch = env.torch
        self.n_total, self.vocab = n_total, vocab
        self.lo, self.hi = env.sharded.shard_range(n_total, env.rank, env.world)
        self.emb, self.post = make_shard(env.synth, env.device, self.lo, self.hi, env.rank, vocab)
        nd, n_docs_total, self.avgdl = env.sharded.global_bm25_stats(
            self.post["nd"], int(self.post["doc_len"].to(torch.int64).sum()), self.hi - self.lo)
        self.nd_global = nd.cpu().numpy()
        self.idf = env.synth.idf_from_counts(n_docs_total, self.nd_global, EPS)
        self.dense = env.engine.DenseIndex(self.emb, borrow=True)
        self.dense.set_shadow(shadow)
        self.bm25 = env.engine.Bm25Index(
            self.post["term_ptr"], self.post["post_doc"], self.post["post_tf"], self.post["doc_len"],
            self.idf, K1, B_PARAM, self.avgdl, n_terms=vocab, n_docs=self.hi - self.lo)
        self.shard = (env.sharded.ShardedHybrid(self.dense, self.bm25, row_base=self.lo)
                      if env.world > 1 else None)

    @property
    def rows(self) -> int:
        return self.hi - self.lo


class QueryBatch:
    def __init__(self, env, batch: int, vocab: int):
        torch = env.torch
        self.B = batch
        self.q_host, self.t_host, self.off_host = make_queries(env.synth, batch, vocab)
        self.q_pin = torch.from_numpy(self.q_host).pin_memory()
        self.t_pin = torch.from_numpy(self.t_host.reshape(-1).copy()).pin_memory()
        self.off_pin = torch.from_numpy(self.off_host).pin_memory()
        dev = env.device
        self.q_dev, self.t_dev, self.off_dev = (self.q_pin.to(dev), self.t_pin.to(dev),
                                                self.off_pin.to(dev))
        self.out_ids = torch.empty((batch, TOPK), dtype=torch.int32, device=dev)
        self.out_scores = torch.empty((batch, TOPK), dtype=torch.float64, device=dev)
        self.out_counts = torch.empty((batch,), dtype=torch.int32, device=dev)
        self.h_ids = torch.empty((batch, TOPK), dtype=torch.int32).pin_memory()
        self.h_scores = torch.empty((batch, TOPK), dtype=torch.float64).pin_memory()
        self.h_counts = torch.empty((batch,), dtype=torch.int32).pin_memory()


def barrier(env):
    env.torch.cuda.synchronize()
    if env.world > 1:
        env.dist.barrier()
    env.torch.cuda.synchronize()


def timed(env, fn, steps: int) -> float:
    """ms for `steps` calls: CUDA events on the launching stream, barrier + synchronize on both
    sides, max over ranks."""
    torch = env.torch
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(env)
    start.record()
    for _ in range(steps):
        fn()
    stop.record()
    barrier(env)
    ms = torch.tensor([start.elapsed_time(stop)], device=env.device, dtype=torch.float64)
    if env.world > 1:
        env.dist.all_reduce(ms, op=env.dist.ReduceOp.MAX)
    return float(ms)


def step_fns(env, w: Workload, qb: QueryBatch):
    """(device-resident step, host-buffer step) for batches of b <= qb.B queries."""
    native, engine, ctx, torch = env.native, env.engine, env.ctx, env.torch

    def step_device(b=qb.B):
        if w.shard is not None:
            return w.shard.search(qb.q_dev[:b], qb.t_dev, qb.off_dev, b, TOPK, W_DENSE, W_BM25,
                                  WRRF_K, TOPK)
        native.call("anr_hybrid_search", ctx.handle, w.dense.handle, w.bm25.handle,
                    qb.q_dev.data_ptr(), qb.t_dev.data_ptr(), qb.off_dev.data_ptr(), b, TOPK, TOPK,
                    None, None, None, 0, W_DENSE, W_BM25, WRRF_K, TOPK, qb.out_ids.data_ptr(),
                    qb.out_scores.data_ptr(), qb.out_counts.data_ptr(), None, None, None, None,
                    engine.torch_stream_ptr())
        return qb.out_ids, qb.out_scores, qb.out_counts

    def step_e2e(b=qb.B):
        """Host buffers in, host buffers out (the call synchronises before returning)."""
        if w.shard is not None:
            qd = qb.q_pin[:b].to(env.device, non_blocking=True)
            td = qb.t_pin.to(env.device, non_blocking=True)
            od = qb.off_pin.to(env.device, non_blocking=True)
            ids, scores, counts = w.shard.search(qd, td, od, b, TOPK, W_DENSE, W_BM25, WRRF_K, TOPK)
            qb.h_ids[:b].copy_(ids, non_blocking=True)
            qb.h_scores[:b].copy_(scores, non_blocking=True)
            qb.h_counts[:b].copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return
        native.call("anr_hybrid_search", ctx.handle, w.dense.handle, w.bm25.handle,
                    qb.q_pin.data_ptr(), qb.t_pin.data_ptr(), qb.off_pin.data_ptr(), b, TOPK, TOPK,
                    None, None, None, 0, W_DENSE, W_BM25, WRRF_K, TOPK, qb.h_ids.data_ptr(),
                    qb.h_scores.data_ptr(), qb.h_counts.data_ptr(), None, None, None, None,
                    engine.torch_stream_ptr())
    return step_device, step_e2e


def profile_reset(env):
    for kind in (0, 1, 2):
        env.native.call("anr_ctx_profile_read", env.ctx.handle, kind, None, None)


def profile_read(env, kind: int):
    ms, n = C.c_double(), C.c_int64()
    env.native.call("anr_ctx_profile_read", env.ctx.handle, kind, C.byref(ms), C.byref(n))
    return ms.value, n.value


def full_lists_single_gpu(env, w: Workload, qb: QueryBatch):
    """The whole batch through the same kernels as the timed steps, with the per-retriever lists."""
    return env.engine.hybrid_search(w.dense, w.bm25, qb.q_host,
                                    [list(map(int, t)) for t in qb.t_host], TOPK, TOPK, W_DENSE,
                                    W_BM25, WRRF_K, TOPK, want_lists=True)


def full_lists_sharded(env, w: Workload, qb: QueryBatch):
    """N > 1: the all-gathered keys of the real NCCL path, decoded on the host: the merged dense
    and BM25 lists (global ids, the scores the GPUs reported) + the fused output."""
    torch = env.torch
    ids, scores, counts = w.shard.search(qb.q_dev, qb.t_dev, qb.off_dev, qb.B, TOPK, W_DENSE, W_BM25,
                                         WRRF_K, TOPK)
    torch.cuda.synchronize()
    gathered = w.shard._buffers(qb.B, TOPK, TOPK)["gathered"].cpu().numpy().view(np.uint64)
    out = dict(ids=ids.cpu().numpy(), scores=scores.cpu().numpy(), counts=counts.cpu().numpy())
    for plane, name in ((0, "dense"), (1, "bm25")):
        keys = gathered[:, plane].transpose(1, 0, 2).reshape(qb.B, -1)       # [B, world * k]
        keys = np.sort(keys, axis=1)[:, ::-1][:, :TOPK]                      # a larger key is a better hit
        sc, idv = decode_keys(keys)
        out["dense_rows" if name == "dense" else "bm25_ids"] = idv
        out[name + "_scores"] = sc
    return out


def decode_keys(keys: np.ndarray):
    """Sortable 64-bit keys (anr_common.cuh: orderable(score) << 32 | ~id) -> (fp32 scores, ids)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    ordered = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(ordered & np.uint32(0x80000000), ordered & np.uint32(0x7fffffff), ~ordered)
    scores = np.ascontiguousarray(bits, dtype=np.uint32).view(np.float32)
    ids = (np.uint64(0xffffffff) - (keys & np.uint64(0xffffffff))).astype(np.int64)
    return scores, ids


def run_headline(env, args, peaks, sampler):
    torch, native, engine = env.torch, env.native, env.engine
    rank, world = env.rank, env.world
    shadow = args.dense_operands == "bf16"
    w = Workload(env, args.chunks, args.vocab, shadow)
    qb = QueryBatch(env, args.batch, args.vocab)
    B = qb.B
    step_device, step_e2e = step_fns(env, w, qb)

    mark = sampler.mark() if sampler else 0
    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    # the headline blocks run as a caller's steps do; the kernel durations of `roofline` come from
    # further blocks of the same steps, right after, with the library's event hooks on (six event
    # records per step, which also sit between kernels that otherwise launch programmatically)
    block_ms = [timed(env, step_device, args.steps) for _ in range(max(args.blocks, 1))]
    pdl_mask = int(os.environ.get("ANR_PDL", "7"))
    native.call("anr_set_option", b"pdl", 0)      # events between kernels need full launch boundaries
    native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
    profile_reset(env)
    prof_ms = [timed(env, step_device, args.steps) for _ in range(3)]
    scan_ms, scan_n = profile_read(env, 0)
    bm_ms, bm_n = profile_read(env, 1)
    pass_ms, pass_n = profile_read(env, 2)
    native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
    native.call("anr_set_option", b"pdl", pdl_mask)
    ms_dev = statistics.median(block_ms)          # one block of exactly `steps` steps
    total_ms = sum(prof_ms)                       # the blocks the kernel durations belong to
    e2e_ms = [timed(env, step_e2e, args.steps) for _ in range(3)]
    ms_e2e = statistics.median(e2e_ms)

    # ---- N > 1: the three parts of a sharded step, each timed alone ---------------------------
    multi = None
    if world > 1:
        ph = w.shard.phases(qb.q_dev, qb.t_dev, qb.off_dev, B, TOPK, W_DENSE, W_BM25, WRRF_K, TOPK)
        for fn in ph.values():
            for _ in range(3):
                fn()
        multi = {"local_ms": timed(env, ph["local"], args.steps) / args.steps,
                 "allgather_ms": timed(env, ph["exchange"], args.steps) / args.steps,
                 "merge_fuse_ms": timed(env, ph["merge_fuse"], args.steps) / args.steps,
                 "allgather_bytes_per_rank": 2 * B * TOPK * 8,
                 "what": "the three parts of one sharded step, each enqueued back to back "
                         "`steps` times and timed alone (max over ranks); in a step they follow "
                         "each other on one stream"}

    # ---- filtered variant: a "CG,NG"-style source filter keeping ~2/3 of the rows,
    #      passed as row / document bit masks (search_engine.py:36-55, :221-231) ----------------
    ms_filtered = None
    if world == 1:
        rng = np.random.default_rng(99)
        keep = rng.random(w.rows) < (2.0 / 3.0)
        words = torch.from_numpy(engine.pack_mask(keep).view(np.int32)).to(env.device)

        def step_filtered():
            native.call("anr_hybrid_search", env.ctx.handle, w.dense.handle, w.bm25.handle,
                        qb.q_dev.data_ptr(), qb.t_dev.data_ptr(), qb.off_dev.data_ptr(), B, TOPK, TOPK,
                        words.data_ptr(), words.data_ptr(), None, 0, W_DENSE, W_BM25, WRRF_K, TOPK,
                        qb.out_ids.data_ptr(), qb.out_scores.data_ptr(), qb.out_counts.data_ptr(),
                        None, None, None, None, engine.torch_stream_ptr())
        for _ in range(3):
            step_filtered()
        ms_filtered = timed(env, step_filtered, args.steps)

    # ---- "content-word" variant: query terms resampled within ranks > 27, since real
    #      queries carry no stop-words (preprocess_bm25.py:41-46 removes them) ----------------------
    ms_content = None
    if world == 1:
        t_cw = env.synth.zipf_queries(B, N_TERMS, args.vocab, ZIPF_S, seed=T_SEED + 1, skip_head=27)
        t_cw_dev = torch.from_numpy(t_cw.reshape(-1).copy()).to(env.device)

        def step_content():
            native.call("anr_hybrid_search", env.ctx.handle, w.dense.handle, w.bm25.handle,
                        qb.q_dev.data_ptr(), t_cw_dev.data_ptr(), qb.off_dev.data_ptr(), B, TOPK, TOPK,
                        None, None, None, 0, W_DENSE, W_BM25, WRRF_K, TOPK,
                        qb.out_ids.data_ptr(), qb.out_scores.data_ptr(), qb.out_counts.data_ptr(),
                        None, None, None, None, engine.torch_stream_ptr())
        for _ in range(3):
            step_content()
        ms_content = timed(env, step_content, args.steps)

    # ---- batch-1 latency through the C ABI with host buffers (p50 of wall-clock per call) -----
    lat = []
    if world == 1:
        for i in range(args.latency_iters + 20):
            t0 = time.perf_counter()
            step_e2e(1)
            if i >= 20:
                lat.append(1e3 * (time.perf_counter() - t0))
    n_b1 = max(args.steps, 50)
    ms_b1_dev = timed(env, lambda: step_device(1), n_b1) / n_b1
    # keep the GPU under the same load until a few clock samples exist (a fixed count derived from
    # the max-reduced step time: every rank runs the same number of collectives)
    for _ in range(int(min(2000, max(10, 500.0 / max(ms_dev / args.steps, 1e-3))))):
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.summary(mark) if sampler else None
    ms_b1_scan, lat_scan = None, []
    if shadow:   # batch-1 again through the fp32 CUDA-core scan (north_star kernel (1))
        w.dense.set_shadow(False)
        for _ in range(5):
            step_device(1)
        ms_b1_scan = timed(env, lambda: step_device(1), n_b1) / n_b1
        if world == 1:
            for i in range(args.latency_iters + 20):
                t0 = time.perf_counter()
                step_e2e(1)
                if i >= 20:
                    lat_scan.append(1e3 * (time.perf_counter() - t0))
        w.dense.set_shadow(True)
        step_device()

    # ---- the whole batch with its per-retriever lists (what the parity check reads) -----------
    got = full_lists_sharded(env, w, qb) if world > 1 else (
        full_lists_single_gpu(env, w, qb) if rank == 0 else None)
    # queries of the benchmark batch that left the fast paths (bf16 nomination margin -> exact fp32
    # rescan; candidate-driven BM25 -> exhaustive scan): both reruns cost several times the step
    reruns = None
    if world == 1:
        d_rr, b_rr = env.ctx.last_rerun()
        reruns = {"dense_queries": d_rr, "bm25_queries": b_rr, "of": B,
                  "what": "queries of the benchmark batch rerun exactly on the device (-1: no such path)"}

    # ---- rank 0 alone from here: rooflines, graph replay, CPU port + parity ---------------------
    res = None
    if rank == 0:
        res = headline_report(env, args, peaks, w, qb, got, dict(
            block_ms=block_ms, prof_ms=prof_ms, ms_dev=ms_dev, total_ms=total_ms, ms_e2e=ms_e2e,
            e2e_ms=e2e_ms,
            scan=(scan_ms, scan_n), bm=(bm_ms, bm_n), tc_pass=(pass_ms, pass_n), multi=multi,
            ms_filtered=ms_filtered, ms_content=ms_content, lat=lat, ms_b1_dev=ms_b1_dev, ms_b1_scan=ms_b1_scan,
            lat_scan=lat_scan, clocks=clocks, shadow=shadow, reruns=reruns))
    del w, qb
    torch.cuda.empty_cache()
    return res


def headline_report(env, args, peaks, w, qb, got, t):
    torch, native, engine = env.torch, env.native, env.engine
    world, B = env.world, qb.B
    peak = peaks["hbm"]
    shadow = t["shadow"]
    rows_local = w.rows
    scan_ms, scan_n = t["scan"]
    bm_ms, bm_n = t["bm"]
    pass_ms, pass_n = t["tc_pass"]
    total_ms, ms_dev = t["total_ms"], t["ms_dev"]
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_tbl = {}
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic_tbl = json.load(fh).get("bytes_per_launch", {})
    # anr_dense_gemm.cu takes batches > 32, and every batch when a bf16 shadow exists
    gemm = (B > 32 or shadow) and rows_local >= 80_000
    tile_q = 64 if B <= 64 else (128 if B <= 128 else 256)
    dense_kernel = (f"dense_gemm_kernel<{tile_q}, {'bf16' if shadow else 'tf32'}>" if gemm else
                    (("dense_tc_pair_kernel" if B > 32 else "dense_tc_kernel<0>") if B > 8
                     else "dense_scan_kernel<1, 4, 0>"))
    at_default = args.chunks == 1_000_000 and world == 1 and B == 64
    scan_bytes = rows_local * D * (2 if (shadow and gemm) else 4)
    scan_avg_ms = scan_ms / max(scan_n, 1)
    achieved = scan_bytes / (scan_avg_ms * 1e-3) / 1e9 if scan_avg_ms > 0 else 0.0
    bm_avg_ms = bm_ms / max(bm_n, 1)
    nd_local = w.post["nd"].cpu().numpy()
    bm_bytes = 8 * int(nd_local[qb.t_host.reshape(-1)].astype(np.int64).sum())
    traffic = traffic_tbl.get(dense_kernel + " co-resident") if at_default else None
    bm_traffic = traffic_tbl.get("bm25 candidate-driven top-k, batch 64") if at_default else None
    dense_roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                  "frac": achieved / peak, "traffic": traffic, "peak_source": peaks["source"],
                  "kernel": dense_kernel, "bytes_per_launch": scan_bytes,
                  "avg_launch_ms": scan_avg_ms, "launches": int(scan_n),
                  "share_of_step": scan_ms / total_ms if total_ms else None,
                  "frac_traffic": (traffic / (scan_avg_ms * 1e-3) / 1e9 / peak
                                   if traffic and scan_avg_ms > 0 else None),
                  "co_resident": "inside the step this kernel shares every SM with the BM25 kernels "
                                 "(its ring leaves them 56 KB of shared memory): avg_launch_ms / "
                                 "achieved / frac are in-step; alone_* is the same batch through a "
                                 "dense-only call"}
    bm_roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peaks["source"],
               "kernel": "BM25 candidate-driven top-k (ms_plan / ms_stage1 / ms_theta / ms_stage2 / "
                         "ms_final kernels, anr_bm25_ms.cu; one `launch` = the whole chain)",
               "bytes_per_launch": bm_bytes,
               "bytes_definition": "UNPRUNED algorithmic figure, 8 B x sum of df over every query-"
                                   "term occurrence (SURVEY 8d): what an exhaustive scan would read. "
                                   "The path touches `traffic` bytes and is bound by the latency "
                                   "of dependent random DRAM accesses, not by bandwidth: frac says "
                                   "how fast an exhaustive scan would have to be to keep up, "
                                   "frac_traffic is the DRAM share it actually uses",
               "traffic": bm_traffic, "launches": int(bm_n), "in_step_ms": bm_avg_ms,
               "in_step_share": bm_ms / total_ms if total_ms else None}
    if world == 1:
        # both kernels again, each ALONE (same batch, same kernels, nothing beside them)
        k_sc = torch.empty((B, TOPK), dtype=torch.float32, device=env.device)
        k_id = torch.empty((B, TOPK), dtype=torch.int32, device=env.device)
        k_ct = torch.empty((B,), dtype=torch.int32, device=env.device)

        def bm25_only():
            native.call("anr_bm25_search", env.ctx.handle, w.bm25.handle, qb.t_dev.data_ptr(),
                        qb.off_dev.data_ptr(), B, TOPK, None, None, 0, k_sc.data_ptr(),
                        k_id.data_ptr(), k_ct.data_ptr(), engine.torch_stream_ptr())

        def dense_only():
            native.call("anr_dense_search", env.ctx.handle, w.dense.handle, qb.q_dev.data_ptr(), B,
                        TOPK, None, 0, k_sc.data_ptr(), k_id.data_ptr(), k_ct.data_ptr(),
                        engine.torch_stream_ptr())
        for fn, kind, roof, nbytes in ((bm25_only, 1, bm_roof, bm_bytes),
                                       (dense_only, 0, dense_roof, scan_bytes)):
            native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
            for _ in range(3):
                fn()
            profile_reset(env)
            for _ in range(10):
                fn()
            a_ms, a_n = profile_read(env, kind)
            native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
            alone = a_ms / max(a_n, 1)
            roof["alone_ms"] = alone
            roof["alone_achieved"] = nbytes / (alone * 1e-3) / 1e9 if alone > 0 else None
            roof["alone_frac"] = roof["alone_achieved"] / peak if alone > 0 else None
    # BM25's headline figures are those of the launch timed alone (its in-step event time includes
    # sharing the SMs with the dense kernel); the in-step time is kept beside them
    bm_time = bm_roof.get("alone_ms") or bm_avg_ms
    bm_roof["avg_launch_ms"] = bm_time
    bm_roof["achieved"] = bm_bytes / (bm_time * 1e-3) / 1e9 if bm_time > 0 else 0.0
    bm_roof["frac"] = bm_roof["achieved"] / peak
    bm_roof["frac_traffic"] = (bm_traffic / (bm_time * 1e-3) / 1e9 / peak
                               if bm_traffic and bm_time > 0 else None)
    bm_roof["share_of_step"] = bm_time * max(bm_n, 1) / total_ms if total_ms else None
    dominant = dense_roof if scan_ms >= bm_time * max(bm_n, 1) else bm_roof

    # ---- the same step replayed from a CUDA graph (graph.HybridGraph) ----------------
    graph_rec = None
    if world == 1:
        try:
            graph_mod = importlib.import_module("a-nice-rag_b200.graph")
            step_device, _ = step_fns(env, w, qb)
            graph_rec = {}
            n_it = max(args.steps, 50)
            for b_g in sorted({1, B}):
                hg = graph_mod.HybridGraph(w.dense, w.bm25, b_g, N_TERMS * b_g, TOPK, TOPK, W_DENSE,
                                           W_BM25, WRRF_K, TOPK)
                hg.load(qb.q_dev[:b_g], qb.t_dev[:N_TERMS * b_g], qb.off_dev[:b_g + 1])
                for _ in range(5):
                    hg.replay()
                step_device(b_g)
                torch.cuda.synchronize()
                same = bool(torch.equal(hg.ids, qb.out_ids[:b_g]) and
                            torch.equal(hg.scores, qb.out_scores[:b_g]) and
                            torch.equal(hg.counts, qb.out_counts[:b_g]))
                ms_g = timed(env, hg.replay, n_it) / n_it
                graph_rec[f"batch{b_g}"] = {"replay_ms": ms_g, "queries_per_s": 1e3 * b_g / ms_g,
                                            "identical_to_eager": same}
                del hg
        except Exception as exc:   # never lose the headline line to the optional replay leg
            graph_rec = {"error": repr(exc)[:300]}

    # ---- where the parts of ONE step start and end (events at fixed marks inside the call) -----
    timeline = None
    if world == 1:
        try:
            names = ["step_begin", "bm25_begin", "bm25_topk_end", "dense_pass_begin",
                     "dense_main_begin", "dense_main_end", "dense_rescore_end", "dense_pass_end",
                     "bm25_rerun_begin", "bm25_rerun_begin_", "bm25_end", "fusion_end",
                     "bm25_plan_end", "bm25_stage1_end", "bm25_theta_end", "bm25_stage2_end"]
            step_device, _ = step_fns(env, w, qb)
            native.call("anr_ctx_timeline_enable", env.ctx.handle, 1)
            runs = []
            for _ in range(9):
                torch.cuda.synchronize()
                step_device()
                marks = (C.c_double * len(names))()
                native.call("anr_ctx_timeline_read", env.ctx.handle, marks)
                runs.append(list(marks))
            native.call("anr_ctx_timeline_enable", env.ctx.handle, 0)
            med = [statistics.median(r[i] for r in runs[2:]) for i in range(len(names))]
            timeline = {"unit": "ms from step_begin, median of 7 single steps (one step in flight, "
                                "the stream idle before it)",
                        **{n: round(v, 4) for n, v in zip(names, med) if not n.endswith("_")}}
        except Exception as exc:
            timeline = {"error": repr(exc)[:200]}

    # ---- two batches in flight: the latency-bound ends of one step (sample passes, thresholds,
    #      rescoring, merges, fusion) run under the main kernels of the other.  `two_in_flight`:
    #      inputs resident, two captured steps alternating on two streams; `e2e_pipelined`:
    #      graph.HybridPipeline, host buffers in and out (H2D / D2H inside), depth 2. -------------
    pipe_rec = None
    if world == 1:
        try:
            graph_mod = importlib.import_module("a-nice-rag_b200.graph")
            step_device, _ = step_fns(env, w, qb)
            step_device()
            torch.cuda.synchronize()
            caps = [graph_mod.HybridGraph(w.dense, w.bm25, B, N_TERMS * B, TOPK, TOPK, W_DENSE, W_BM25,
                                          WRRF_K, TOPK) for _ in range(2)]
            streams = [torch.cuda.Stream(env.device) for _ in range(2)]
            for hg in caps:
                hg.load(qb.q_dev, qb.t_dev, qb.off_dev)
                hg.replay()
            torch.cuda.synchronize()
            same = all(bool(torch.equal(hg.ids, qb.out_ids) and torch.equal(hg.scores, qb.out_scores))
                       for hg in caps)

            def two_steps():
                cur = torch.cuda.current_stream()
                for hg, st in zip(caps, streams):
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        hg.replay()
                for st in streams:
                    cur.wait_stream(st)
            n_it = max(args.steps, 50) // 2
            for _ in range(5):
                two_steps()
            ms2 = timed(env, two_steps, n_it) / n_it / 2
            pipe_rec = {"two_in_flight": {"ms_per_step": ms2, "queries_per_s": 1e3 * B / ms2,
                                          "identical_to_eager": same,
                                          "what": "device-resident inputs, two captured steps "
                                                  "alternating on two streams"}}
            del caps
            pipe = graph_mod.HybridPipeline(w.dense, w.bm25, B, N_TERMS * B, TOPK, TOPK, W_DENSE,
                                            W_BM25, WRRF_K, TOPK, depth=2)
            t_flat = qb.t_host.reshape(-1)

            def run_pipe(n_steps):
                last = None
                for i in range(n_steps):
                    if i >= 2:
                        last = pipe.collect()
                    pipe.submit(qb.q_host, t_flat, qb.off_host)
                for _ in range(min(2, n_steps)):
                    last = pipe.collect()
                return last
            run_pipe(6)
            last = run_pipe(3)
            same_p = bool(np.array_equal(last[0], qb.out_ids.cpu().numpy()) and
                          np.array_equal(last[1], qb.out_scores.cpu().numpy()))
            n_it = max(args.steps, 50)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_pipe(n_it)
            dt = time.perf_counter() - t0
            pipe_rec["e2e_pipelined"] = {
                "value": B * n_it / dt, "unit": "queries/s", "ms_per_step": 1e3 * dt / n_it, "depth": 2,
                "identical_to_eager": same_p,
                "timing": "host wall clock over the whole loop (every result read on the host)"}
            del pipe
        except Exception as exc:
            pipe_rec = dict(pipe_rec or {}, error=repr(exc)[:300])

    # ---- CPU baseline + parity of EVERY query of this very batch --------------------------------
    cpu, checked, parity_error = None, 0, None
    if not args.no_cpu_baseline:
        cores = set_cpu_threads()
        shards = []
        for r in range(world):   # rank 0 re-creates the other ranks' shards (seeded generators)
            lo, hi = env.sharded.shard_range(args.chunks, r, world)
            emb_r, post_r = (w.emb, w.post) if r == env.rank else make_shard(
                env.synth, env.device, lo, hi, r, args.vocab)
            post_host = {k: v.cpu().numpy() for k, v in post_r.items()}
            shards.append((lo, emb_r.cpu().numpy(), cpu_index(post_host, w.idf, w.avgdl)))
            del emb_r, post_r
        nq_cpu = min(args.cpu_queries, B)
        which = list(range(nq_cpu))
        cpu_port_run(shards, qb.q_host, qb.t_host, [0])                      # warm caches / BLAS
        dt, results = cpu_port_run(shards, qb.q_host, qb.t_host, which)
        cpu = {"value": nq_cpu / dt, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{nq_cpu} of the batch's {B} queries over the full {args.chunks}-chunk "
                         "corpus; numpy BLAS dot + CSR BM25 + Python RRF on a pre-stacked matrix "
                         "(the reference's per-query np.stack and pandas work NOT included)"}
        try:
            cpu["stock_path_config0"] = stock_path_config0()
        except Exception as exc:
            cpu["stock_path_config0"] = {"error": repr(exc)[:200]}
        try:
            checked = check_against_cpu(results, got, shards, qb.q_host)
        except AssertionError as exc:
            parity_error = str(exc)[:500]
        del shards

    steps = args.steps
    line = {
        "metric": METRIC, "value": B * steps / (ms_dev * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "weak" if args.chunks_per_gpu else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args),
        "shard": {"rows_local": rows_local, "bm25_postings_local": w.bm25.n_postings},
        "blocks": {"n": len(t["block_ms"]), "steps_per_block": steps,
                   "what": "value / ms_per_step are the median block; every block is exactly "
                           "`steps` steps between two barriers, max over ranks",
                   "ms_per_step_min": min(t["block_ms"]) / steps,
                   "ms_per_step_max": max(t["block_ms"]) / steps,
                   "ms_per_step_all": [m / steps for m in t["block_ms"]],
                   "profiled": {"n": len(t["prof_ms"]),
                                "ms_per_step": statistics.median(t["prof_ms"]) / steps,
                                "what": "further blocks of the same steps with the library's event "
                                        "hooks on and every launch fully serialised (an event "
                                        "between two kernels needs a full launch boundary; the "
                                        "headline blocks launch their chains programmatically, "
                                        "anr_set_option \"pdl\"): the kernel durations and shares "
                                        "of `roofline` / `roofline_other` / `dense_tc_pass` are "
                                        "measured in these"}},
        "e2e": e2e_record(B, steps, t, qb, pipe_rec),
        # kernels of this repo launched per step.  GEMM path: query -> bf16, sample pass,
        # thresholds, GEMM, rescore, flag compaction, flagged rescan, its merge (8; 7 with tf32
        # operands); BM25 candidate-driven chain: plan, stage 1, theta, stage 2, ranking + the
        # rerun pair for flagged queries (7); WRRF (1; sharded: merges + fusion are one launch)
        "gpu_launches": int(steps * len(t["block_ms"]) * (
            ((8 if shadow else 7) if gemm else ((8 if B > 32 else 7) if B > 8 else 2)) + 7 + 1)),
        "roofline": dominant,
        "roofline_other": bm_roof if dominant is dense_roof else dense_roof,
        "step_traffic": ({"dram_bytes": traffic + bm_traffic, "ms_at_peak": (traffic + bm_traffic) / peak / 1e6,
                          "frac_of_peak": (traffic + bm_traffic) / peak / 1e6 / (ms_dev / steps)}
                         if traffic and bm_traffic else None),
        "dense_operands": ("bf16 shadow copy" if shadow else "fp32 words as tf32") if (B > 8 or gemm) else "fp32",
        "dense_tc_pass": {"what": "sample pre-pass + threshold + scan + exact rescoring",
                          "avg_ms": pass_ms / max(pass_n, 1), "passes": int(pass_n),
                          "share_of_step": pass_ms / total_ms if total_ms else None},
        "batch1": {"device_ms": t["ms_b1_dev"],
                   "device_qps": 1e3 / t["ms_b1_dev"] if t["ms_b1_dev"] else None,
                   "e2e_p50_ms": statistics.median(t["lat"]) if t["lat"] else None,
                   "e2e_p95_ms": (sorted(t["lat"])[int(0.95 * len(t["lat"]))] if t["lat"] else None),
                   "dense_path": ("bf16 shadow GEMM pass + exact fp32 rescoring" if shadow
                                  else "fp32 CUDA-core scan")},
        "batch1_fp32_scan": ({"device_ms": t["ms_b1_scan"],
                              "e2e_p50_ms": statistics.median(t["lat_scan"]) if t["lat_scan"] else None,
                              "hbm_gbs": rows_local * D * 4 / (t["ms_b1_scan"] * 1e-3) / 1e9,
                              "frac_of_peak": rows_local * D * 4 / (t["ms_b1_scan"] * 1e-3) / 1e9 / peak}
                             if t["ms_b1_scan"] else None),
        "filtered": ({"what": "same step with row / document masks keeping 2/3 of the corpus "
                              "(the reference's source-prefix filter)",
                      "value": B * steps / (t["ms_filtered"] * 1e-3), "unit": "queries/s",
                      "ms_per_step": t["ms_filtered"] / steps} if t["ms_filtered"] else None),
        "content_words": ({"what": "same step, BM25 query terms resampled within Zipf ranks > 27 "
                                   "(SURVEY 8d: real queries carry no stop-words)",
                           "value": B * steps / (t["ms_content"] * 1e-3), "unit": "queries/s",
                           "ms_per_step": t["ms_content"] / steps} if t.get("ms_content") else None),
        "cuda_graph": graph_rec,
        "timeline": timeline,
        "reruns": t.get("reruns"),
        "pipelined": pipe_rec,
        "multi_gpu": t["multi"],
        "clocks": t["clocks"], "parity_checked_queries": checked,
    }
    if parity_error:
        line["parity_error"] = parity_error
    if cpu:
        line["cpu_baseline"] = cpu
    return line


def e2e_record(B, steps, t, qb, pipe_rec):
    """End to end through the public API with HOST buffers.  The throughput API for host batches is
    graph.HybridPipeline (submit / collect, two batches in flight: H2D of the queries, the step,
    D2H of the fused result, every result read on the host) -- that is `value` when it ran and
    returned the eager call's results bit for bit; the one-batch-at-a-time synchronous C-ABI call
    (anr_hybrid_search with host pointers) is kept beside it as `synchronous_call`."""
    sync = {"value": B * steps / (t["ms_e2e"] * 1e-3), "unit": "queries/s",
            "ms_per_step_all": [m / steps for m in t["e2e_ms"]],
            "what": "anr_hybrid_search with pinned host pointers, one batch in flight, CUDA events"}
    rec = {"value": sync["value"], "unit": "queries/s", "path": "synchronous C-ABI call",
           "h2d_bytes_per_step": int(qb.q_pin.numel() * 4 + qb.t_pin.numel() * 4 +
                                     qb.off_pin.numel() * 4),
           "d2h_bytes_per_step": int(qb.h_ids.numel() * 4 + qb.h_scores.numel() * 8 +
                                     qb.h_counts.numel() * 4),
           "synchronous_call": sync}
    piped = (pipe_rec or {}).get("e2e_pipelined")
    if piped and piped.get("identical_to_eager") and piped["value"] > sync["value"]:
        rec.update(value=piped["value"], path="graph.HybridPipeline, depth 2 (host wall clock over "
                                              "the whole loop, every result read on the host)",
                   ms_per_step=piped["ms_per_step"])
    return rec


# ---------------------------------------------------------------------------------------
# extra legs: the other BASELINE configs, bounded
# ---------------------------------------------------------------------------------------
def slab_iter(env, emb, slab_rows: int = 1 << 18):
    """(first row, host fp32 [rows, D]) slabs of a device matrix through one pinned buffer."""
    torch = env.torch
    n = emb.shape[0]
    pin = torch.empty((min(slab_rows, n), emb.shape[1]), dtype=torch.float32).pin_memory()
    for r0 in range(0, n, slab_rows):
        r1 = min(n, r0 + slab_rows)
        pin[:r1 - r0].copy_(emb[r0:r1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        yield r0, pin[:r1 - r0].numpy()


def bm25_term_fetcher(post):
    tp = post["term_ptr"]

    def fetch(t: int):
        lo, hi = int(tp[t]), int(tp[t + 1])
        return post["post_doc"][lo:hi].cpu().numpy(), post["post_tf"][lo:hi].cpu().numpy()
    return fetch


def measure_tf32_peak(env):
    """cuBLAS tf32 throughput the way MEASURED_PEAKS.json measures bf16: 8192^3, best of 6."""
    torch = env.torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((8192, 8192), device=env.device)
        b = torch.randn((8192, 8192), device=env.device)
        best = 1e9
        for _ in range(8):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(a, b)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_big_legs(env, args, peaks, sampler):
    """BASELINE configs[2], configs[3] and the batch-1 10M target on ONE GPU (rank 0, world 1)."""
    from oracle import retrieval, slabs
    torch, native, engine, synth = env.torch, env.native, env.engine, env.synth
    n, vocab, steps = args.big_chunks, args.big_vocab, max(args.leg_steps, 3)
    out = {}
    stream = engine.torch_stream_ptr
    emb = synth.unit_vectors_torch(n, D, EMB_SEED, env.device)
    dense = engine.DenseIndex(emb, borrow=True)
    dense.set_shadow(True)

    # ---- configs[2]: 10M x 1024, batch 1024, dense-only top-100 ------------------------------
    b3, k3 = 1024, 100
    q3_host = synth.unit_vectors(b3, D, seed=Q_SEED)
    q3 = torch.from_numpy(q3_host).to(env.device)
    sc3 = torch.empty((b3, k3), dtype=torch.float32, device=env.device)
    id3 = torch.empty((b3, k3), dtype=torch.int32, device=env.device)
    ct3 = torch.empty((b3,), dtype=torch.int32, device=env.device)

    def dense_step():
        native.call("anr_dense_search", env.ctx.handle, dense.handle, q3.data_ptr(), b3, k3, None, 0,
                    sc3.data_ptr(), id3.data_ptr(), ct3.data_ptr(), stream())
    rec3 = {"workload": f"{n} chunks x {D}-d, batch {b3}, dense-only top-{k3} "
                        "(tcgen05 GEMM nomination + exact fp32 rescoring)"}
    flops = 2.0 * b3 * n * D
    for operands in ("bf16", "tf32"):
        dense.set_shadow(operands == "bf16")
        mark = sampler.mark() if sampler else 0
        for _ in range(3):
            dense_step()
        native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
        profile_reset(env)
        ms = timed(env, dense_step, steps) / steps
        k_ms, k_n = profile_read(env, 0)
        native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
        kernel_ms = k_ms / max(k_n, 1)
        r = {"ms_per_batch": ms, "queries_per_s": 1e3 * b3 / ms, "steps": steps,
             "kernel_ms": kernel_ms, "tflops_kernel": flops / (kernel_ms * 1e-3) / 1e12,
             "tflops_step": flops / (ms * 1e-3) / 1e12,
             "clocks": sampler.summary(mark) if sampler else None}
        if operands == "bf16":
            r["roofline"] = {"bound": "tensor", "achieved": r["tflops_kernel"], "peak": peaks["bf16"],
                             "unit": "TFLOP/s", "frac": r["tflops_kernel"] / peaks["bf16"],
                             "peak_sustained": peaks["bf16_sustained"],
                             "frac_sustained": r["tflops_kernel"] / peaks["bf16_sustained"],
                             "peak_source": peaks["source"], "traffic": None,
                             "kernel": "dense_gemm2_kernel<256, bf16> (cta_group::2)"}
            if k3 == 100:
                rec3["rows_bf16"] = id3.cpu().numpy()
                rec3["scores_bf16"] = sc3.cpu().numpy()
        else:
            tf32_peak = measure_tf32_peak(env)
            r["roofline"] = {"bound": "tensor", "achieved": r["tflops_kernel"], "peak": tf32_peak,
                             "unit": "TFLOP/s", "frac": r["tflops_kernel"] / tf32_peak,
                             "peak_source": "measured in this run: torch.matmul fp32 with "
                                            "allow_tf32, 8192^3, best of 8", "traffic": None,
                             "kernel": "dense_gemm2_kernel<256, tf32> (cta_group::2)"}
            rec3["identical_to_bf16"] = bool(
                np.array_equal(rec3["rows_bf16"], id3.cpu().numpy()) and
                np.array_equal(rec3["scores_bf16"], sc3.cpu().numpy()))
        rec3[operands] = r
    dense.set_shadow(True)
    rows3, scores3 = rec3.pop("rows_bf16"), rec3.pop("scores_bf16")

    # ---- BM25 index over the same 10M documents (V = 500k) --------------------------------------
    post = synth.zipf_postings_torch(n, vocab, ZIPF_S, POST_SEED, env.device)
    nd = post["nd"].cpu().numpy()
    idf = synth.idf_from_counts(n, nd, EPS)
    avgdl = float(post["doc_len"].to(torch.int64).sum()) / n
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                            idf, K1, B_PARAM, avgdl, n_terms=vocab, n_docs=n)

    # ---- configs[3]: BM25-only, batch 256, top-10 --------------------------------------------
    b4 = 256
    tq4 = synth.zipf_queries(b4, N_TERMS, vocab, ZIPF_S, seed=T_SEED)
    t4 = torch.from_numpy(tq4.reshape(-1).copy()).to(env.device)
    o4 = torch.arange(0, (b4 + 1) * N_TERMS, N_TERMS, dtype=torch.int32, device=env.device)
    sc4 = torch.empty((b4, TOPK), dtype=torch.float32, device=env.device)
    id4 = torch.empty((b4, TOPK), dtype=torch.int32, device=env.device)
    ct4 = torch.empty((b4,), dtype=torch.int32, device=env.device)

    def bm25_step():
        native.call("anr_bm25_search", env.ctx.handle, bm25.handle, t4.data_ptr(), o4.data_ptr(), b4,
                    TOPK, None, None, 0, sc4.data_ptr(), id4.data_ptr(), ct4.data_ptr(), stream())
    mark = sampler.mark() if sampler else 0
    for _ in range(3):
        bm25_step()
    native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
    profile_reset(env)
    ms4 = timed(env, bm25_step, steps) / steps
    k_ms, k_n = profile_read(env, 1)
    native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
    kernel4 = k_ms / max(k_n, 1)
    bytes4 = 8 * int(nd[tq4.reshape(-1)].astype(np.int64).sum())
    rec4 = {"workload": f"BM25-only, {n} docs, V={vocab}, Zipf {ZIPF_S}, {N_TERMS}-term queries, "
                        f"batch {b4}, top-{TOPK}",
            "postings": bm25.n_postings, "ms_per_batch": ms4, "queries_per_s": 1e3 * b4 / ms4,
            "steps": steps, "sum_df_per_query": bytes4 / 8 / b4,
            "roofline": {"bound": "hbm", "achieved": bytes4 / (kernel4 * 1e-3) / 1e9,
                         "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": bytes4 / (kernel4 * 1e-3) / 1e9 / peaks["hbm"],
                         "bytes_per_launch": bytes4, "avg_launch_ms": kernel4,
                         "bytes_definition": "UNPRUNED algorithmic figure (8 B x sum of df): what an "
                                             "exhaustive scan would read; the candidate-driven "
                                             "path touches a small part of it (latency-bound "
                                             "lookups), so frac exceeds 1",
                         "traffic": None, "peak_source": peaks["source"],
                         "kernel": "BM25 candidate-driven top-k (anr_bm25_ms.cu, the whole chain)"},
            "clocks": sampler.summary(mark) if sampler else None}

    # ---- batch-1 hybrid at 10M (the >= 70 % of HBM roofline target) -----------------------------
    qb = QueryBatch(env, 1, vocab)
    rec1 = {"workload": workload_name(n, vocab) + ", batch 1"}

    def hybrid1(want=False):
        return engine.hybrid_search(dense, bm25, qb.q_host, [list(map(int, qb.t_host[0]))], TOPK,
                                    TOPK, W_DENSE, W_BM25, WRRF_K, TOPK, want_lists=True)

    def hybrid1_dev():
        native.call("anr_hybrid_search", env.ctx.handle, dense.handle, bm25.handle,
                    qb.q_dev.data_ptr(), qb.t_dev.data_ptr(), qb.off_dev.data_ptr(), 1, TOPK, TOPK,
                    None, None, None, 0, W_DENSE, W_BM25, WRRF_K, TOPK, qb.out_ids.data_ptr(),
                    qb.out_scores.data_ptr(), qb.out_counts.data_ptr(), None, None, None, None,
                    stream())
    got1 = {}
    for path in ("bf16_shadow", "fp32_scan"):
        dense.set_shadow(path == "bf16_shadow")
        mark = sampler.mark() if sampler else 0
        for _ in range(3):
            hybrid1_dev()
        native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
        profile_reset(env)
        ms = timed(env, hybrid1_dev, steps) / steps
        k_ms, k_n = profile_read(env, 0)
        native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
        nbytes = n * D * (2 if path == "bf16_shadow" else 4)
        kernel_ms = k_ms / max(k_n, 1)
        rec1[path] = {"device_ms": ms, "steps": steps, "dense_kernel_ms": kernel_ms,
                      "bytes_per_step": nbytes,
                      "hbm_gbs_step": nbytes / (ms * 1e-3) / 1e9,
                      "frac_of_hbm_peak_step": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm"],
                      "frac_of_hbm_peak_kernel": (nbytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm"]
                                                  if kernel_ms > 0 else None),
                      "clocks": sampler.summary(mark) if sampler else None}
        got1[path] = hybrid1()
    dense.set_shadow(True)
    rec1["target"] = ">= 0.70 of the measured HBM peak for the batch-1 hybrid step (north_star)"

    # ---- parity: ONE slab sweep of the corpus serves configs[2] (4 sampled queries, top-100)
    #      and the batch-1 query (= query 0, top-10); BM25 from the postings of the query terms ---
    sample3 = [0, 341, 682, 1023]
    try:
        want_ids, want_sc = slabs.dense_topk_slabs(slab_iter(env, emb), q3_host[sample3], k3)

        def score_of_factory(qv):
            return lambda row: float(np.dot(emb[row].cpu().numpy(), qv))
        for j, q in enumerate(sample3):
            slabs.assert_topk_matches(rows3[q], scores3[q], want_ids[j], want_sc[j],
                                      score_of_factory(q3_host[q]), what=f"config3 q{q}")
        rec3["parity_checked_queries"] = len(sample3)
        fetch = bm25_term_fetcher(post)
        doc_len_host = post["doc_len"].cpu().numpy()
        idf_of = lambda t: float(idf[t])   # noqa: E731
        # batch-1 hybrid: both dense paths, same lists
        b_all = slabs.bm25_scores_subindex(fetch, qb.t_host[0], doc_len_host, idf_of, avgdl, K1, B_PARAM)
        b_docs = retrieval.bm25_topk(b_all, TOPK)
        for path, g in got1.items():
            slabs.assert_topk_matches(g["dense_rows"][0], g["dense_scores"][0], want_ids[0][:TOPK],
                                      want_sc[0][:TOPK], score_of_factory(q3_host[0]),
                                      what=f"batch1_10M dense {path}")
            retrieval.assert_ranking_matches(g["bm25_ids"][0], g["bm25_scores"][0], b_docs,
                                             b_all[b_docs], all_scores=b_all, rtol=1e-5, atol=1e-6,
                                             what=f"batch1_10M bm25 {path}")
            from oracle import pipeline
            c = int(g["counts"][0])
            pipeline.check_fused(g["ids"][0, :c], g["scores"][0, :c], g["dense_rows"][0],
                                 g["bm25_ids"][0], (W_DENSE, W_BM25), WRRF_K, TOPK)
        rec1["parity_checked_queries"] = 1
        # configs[3]: 4 sampled queries of the batch
        ids4, scs4 = id4.cpu().numpy(), sc4.cpu().numpy()
        sample4 = [0, 85, 170, 255]
        for q in sample4:
            b_all = slabs.bm25_scores_subindex(fetch, tq4[q], doc_len_host, idf_of, avgdl, K1, B_PARAM)
            b_docs = retrieval.bm25_topk(b_all, TOPK)
            retrieval.assert_ranking_matches(ids4[q], scs4[q], b_docs, b_all[b_docs],
                                             all_scores=b_all, rtol=1e-5, atol=1e-6,
                                             what=f"config4 q{q}")
        rec4["parity_checked_queries"] = len(sample4)
    except AssertionError as exc:
        out["parity_error"] = str(exc)[:500]
    out["config3"], out["config4"], out["batch1_10M"] = rec3, rec4, rec1
    del dense, bm25, emb, post
    torch.cuda.empty_cache()
    return out


def run_weak_leg(env, args, peaks, sampler):
    """BASELINE configs[4] protocol (i): 12.5M chunks per GPU (100M on 8 GPUs), hybrid batch 64 and
    batch 1; per-GPU bytes are constant, so the ideal latency is.  Parity: every rank scores the
    sampled queries over ITS shard with the oracle (dense slab-wise, BM25 from the query terms'
    postings, global idf), rank 0 merges those lists and compares with the all-gathered result."""
    from oracle import pipeline, retrieval, slabs
    torch, engine = env.torch, env.engine
    rank, world = env.rank, env.world
    per_gpu, steps = args.weak_chunks_per_gpu, max(args.leg_steps, 3)
    n_total = per_gpu * world
    free = engine.context(env.device.index).info()[2]
    need = per_gpu * D * 6 + per_gpu * 160 * 14        # matrix + shadow + postings (+ build slack)
    ok = torch.tensor([1 if free > need else 0], device=env.device)
    if world > 1:
        env.dist.all_reduce(ok, op=env.dist.ReduceOp.MIN)
    if int(ok) == 0:
        return {"skipped": f"needs ~{need / 1e9:.0f} GB free per GPU, {free / 1e9:.0f} GB available"}
    w = Workload(env, n_total, args.vocab, True)
    rec = {"workload": workload_name(n_total, args.vocab) + f", {per_gpu} chunks per GPU",
           "n_gpus": world, "scaling": "weak", "rows_local": w.rows,
           "bm25_postings_local": w.bm25.n_postings}
    got = None
    for B in (64, 1):
        qb = QueryBatch(env, B, args.vocab) if B == 64 else qb64_view(env, qb64, 1)
        if B == 64:
            qb64 = qb
        step_device, _ = step_fns(env, w, qb)
        mark = sampler.mark() if sampler else 0
        for _ in range(3):
            step_device()
        env.native.call("anr_ctx_profile_enable", env.ctx.handle, 1)
        profile_reset(env)
        ms = timed(env, step_device, steps) / steps
        k_ms, k_n = profile_read(env, 0)
        env.native.call("anr_ctx_profile_enable", env.ctx.handle, 0)
        kernel_ms = k_ms / max(k_n, 1)
        nbytes = w.rows * D * 2       # the bf16 shadow pass reads every local row once
        rec[f"batch{B}"] = {
            "ms_per_step": ms, "queries_per_s": 1e3 * B / ms, "steps": steps,
            "dense_kernel_ms": kernel_ms, "bytes_per_gpu_per_step": nbytes,
            "frac_of_hbm_peak_step": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm"],
            "frac_of_hbm_peak_kernel": (nbytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm"]
                                        if kernel_ms > 0 else None),
            "clocks": sampler.summary(mark) if sampler and rank == 0 else None}
        if B == 64:
            got = (full_lists_sharded(env, w, qb) if world > 1 else
                   full_lists_single_gpu(env, w, qb))
    # ---- parity on sampled queries of the batch-64 result ------------------------------------
    sample = [0, 21, 42, 63]
    try:
        set_cpu_threads()
        ids_l, sc_l = slabs.dense_topk_slabs(slab_iter(env, w.emb), qb64.q_host[sample], TOPK)
        fetch = bm25_term_fetcher(w.post)
        doc_len_host = w.post["doc_len"].cpu().numpy()
        idf_of = lambda t: float(w.idf[t])   # noqa: E731
        local = []
        for j, q in enumerate(sample):
            b_all = slabs.bm25_scores_subindex(fetch, qb64.t_host[q], doc_len_host, idf_of, w.avgdl,
                                               K1, B_PARAM)
            b_docs = retrieval.bm25_topk(b_all, TOPK)
            local.append(dict(dense_ids=ids_l[j] + w.lo, dense_scores=sc_l[j],
                              bm25_ids=b_docs.astype(np.int64) + w.lo, bm25_scores=b_all[b_docs]))
        gathered = [local]
        if world > 1:
            gathered = [None] * world
            env.dist.all_gather_object(gathered, local)
        if rank == 0:
            for j, q in enumerate(sample):
                merged = {}
                for name in ("dense", "bm25"):
                    ids = np.concatenate([g[j][name + "_ids"] for g in gathered])
                    sc = np.concatenate([g[j][name + "_scores"] for g in gathered])
                    # descending score, ties by ascending global id (the shards are in id order)
                    order = np.lexsort((ids, -sc))[:TOPK]
                    merged[name] = (ids[order], sc[order])
                pool = {int(i): float(s) for g in gathered
                        for i, s in zip(g[j]["dense_ids"], g[j]["dense_scores"])}
                slabs.assert_topk_matches(got["dense_rows"][q, :TOPK], got["dense_scores"][q, :TOPK],
                                          merged["dense"][0], merged["dense"][1],
                                          lambda i: pool.get(i, -np.inf), what=f"weak dense q{q}")
                poolb = {int(i): float(s) for g in gathered
                         for i, s in zip(g[j]["bm25_ids"], g[j]["bm25_scores"])}
                slabs.assert_topk_matches(got["bm25_ids"][q, :TOPK], got["bm25_scores"][q, :TOPK],
                                          merged["bm25"][0], merged["bm25"][1],
                                          lambda i: poolb.get(i, -np.inf), what=f"weak bm25 q{q}")
                c = int(got["counts"][q])
                pipeline.check_fused(got["ids"][q, :c], got["scores"][q, :c],
                                     got["dense_rows"][q, :TOPK], got["bm25_ids"][q, :TOPK],
                                     (W_DENSE, W_BM25), WRRF_K, TOPK)
            rec["parity_checked_queries"] = len(sample)
    except AssertionError as exc:
        rec["parity_error"] = str(exc)[:500]
    del w
    torch.cuda.empty_cache()
    return rec


def qb64_view(env, qb64, b: int):
    """A batch of the first b queries of an existing batch (same device buffers)."""
    v = QueryBatch.__new__(QueryBatch)
    v.__dict__.update(qb64.__dict__)
    v.B = b
    return v


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    env = Env()
    env.torch, env.dist, env.rank, env.world = torch, dist, rank, world
    env.device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=env.device)
    pkg = importlib.import_module("a-nice-rag_b200")
    env.engine, env.native = pkg.engine, pkg.native
    env.synth = importlib.import_module("a-nice-rag_b200.synth")
    env.sharded = importlib.import_module("a-nice-rag_b200.sharded")
    env.ctx = env.engine.context(local_rank)
    peaks = load_peaks()
    legs = {x.strip() for x in args.legs.split(",") if x.strip()}
    sampler = ClockSampler(local_rank) if rank == 0 else None

    line = run_headline(env, args, peaks, sampler)
    extra = {}
    # The extra legs are bounded (~30 s on one GPU) but collective at N > 1: if one of them ever
    # hung (a rank lost, a wedged collective) the driver would kill the run and lose the headline
    # with it.  Past the budget rank 0 prints the headline line with what it has and every rank
    # leaves.
    def _give_up():
        if rank == 0:
            line["legs"] = dict(extra, aborted=f"extra legs exceeded --leg-budget-s {args.leg_budget_s}")
            print(json.dumps(line), flush=True)
        os._exit(0)
    watchdog = threading.Timer(args.leg_budget_s, _give_up)
    watchdog.daemon = True
    watchdog.start()
    if "big" in legs and world == 1:
        try:
            extra.update(run_big_legs(env, args, peaks, sampler))
        except Exception as exc:       # never lose the headline line to an extra leg
            extra["big_error"] = repr(exc)[:400]
            torch.cuda.empty_cache()
    if "weak" in legs:
        try:
            weak = run_weak_leg(env, args, peaks, sampler)
            if rank == 0:
                extra["weak"] = weak
        except Exception as exc:
            extra["weak"] = {"error": repr(exc)[:400]}
            torch.cuda.empty_cache()
    watchdog.cancel()
    if sampler:
        sampler.stop()
    if rank == 0:
        line["legs"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.chunks_per_gpu:
        args.chunks = args.chunks_per_gpu * dist_env()[2]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
