#!/usr/bin/env python
"""Headline benchmark: hybrid queries/s (dense top-10 + BM25 top-10 + weighted RRF -> top-10)
over the BASELINE.json configs[1] corpus -- 1M chunks x 1024-d fp32 + a Zipf(1.1) BM25 corpus of
1M documents (V=50k, 8-term queries) -- in batches of 64 queries, plus batch-1 latency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A step = one batch of B hybrid queries.
  value     : queries/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       : the same through the C-ABI with pinned HOST buffers (H2D of the queries and D2H
              of the fused results inside the timed region)
  roofline  : dense scan kernel, algorithmic bytes (rows x D x 4 per pass) / mean launch time,
              against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port (numpy: BLAS dot + CSR BM25 + Python RRF, pre-stacked
              matrix, i.e. WITHOUT the reference's per-query np.stack / pandas overhead) on a
              bounded sample of the same batch, same corpus, on this box's host cores
  cuda_graph : the same batch-1 / batch-B steps replayed from a captured CUDA graph
              (a-nice-rag_b200/graph.py), compared bit for bit with the eager call's output
N > 1: the SAME 1M-chunk corpus is sharded by chunk over the ranks (strong scaling); local
top-k keys are exchanged with one NCCL all-gather and merged + fused on every rank.
--impl reference: only the CPU port is timed (rank 0), none of the CUDA library is loaded.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
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
VOCAB = 50_000
ZIPF_S = 1.1
K1, B_PARAM, EPS = 1.7, 0.83, 0.05
W_DENSE, W_BM25, WRRF_K = 5.0, 1.0, 40.0
TOPK = 10
N_TERMS = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=1_000_000, help="total corpus rows / BM25 docs")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--vocab", type=int, default=VOCAB)
    ap.add_argument("--chunks-per-gpu", type=int, default=0,
                    help="weak scaling: total corpus = this x world size (overrides --chunks)")
    ap.add_argument("--cpu-queries", type=int, default=16, help="queries in the CPU sample")
    ap.add_argument("--latency-iters", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-operands", default="bf16", choices=["bf16", "tf32"],
                    help="operands of the tensor-core nomination pass for batches > 32: a bf16 "
                         "shadow copy of the matrix (half the HBM bytes) or the fp32 words read as "
                         "tf32; results are identical (exact fp32 rescoring)")
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
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


# ---------------------------------------------------------------------------------------
# synthetic workload (device generation; identical bytes are copied to the host for the CPU arm)
# ---------------------------------------------------------------------------------------
def make_queries(batch: int):
    synth = importlib.import_module("a-nice-rag_b200.synth")
    q = synth.unit_vectors(batch, D, seed=4321)
    terms = synth.zipf_queries(batch, N_TERMS, VOCAB, ZIPF_S, seed=2025)   # VOCAB = --vocab
    offsets = np.arange(0, (batch + 1) * N_TERMS, N_TERMS, dtype=np.int32)
    return q, terms, offsets


def make_shard(torch, device, lo: int, hi: int, shard_id: int):
    synth = importlib.import_module("a-nice-rag_b200.synth")
    emb = synth.unit_vectors_torch(hi - lo, D, 1234 + shard_id, device)
    post = synth.zipf_postings_torch(hi - lo, VOCAB, ZIPF_S, 2024 + shard_id, device)
    return emb, post


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port, timed (and used as the checker of the GPU results)
# ---------------------------------------------------------------------------------------
def cpu_port_setup(emb_host, post_host, idf, avgdl):
    from oracle import csr
    return emb_host, csr.CsrIndex(
        term_ptr=post_host["term_ptr"], post_doc=post_host["post_doc"],
        post_tf=post_host["post_tf"], doc_len=post_host["doc_len"].astype(np.int64), idf=idf,
        avgdl=avgdl, k1=K1, b=B_PARAM)


def cpu_port_run(emb, index, queries, terms, n_queries: int):
    """Runs n_queries hybrid queries on the CPU; returns (seconds, results)."""
    from oracle import pipeline
    weights = {"voyage-3-large": W_DENSE, "BM25": W_BM25}
    results = []
    t0 = time.perf_counter()
    for q in range(n_queries):
        results.append(pipeline.hybrid_query(queries[q], emb, index, [int(t) for t in terms[q]],
                                             TOPK, TOPK, weights, WRRF_K, TOPK))
    return time.perf_counter() - t0, results


def check_against_cpu(results, got, queries_n: int):
    """GPU fused output vs the oracle on the sampled queries (tie-aware on the two lists, exact
    on the fusion given the lists).  Returns the number of queries checked."""
    from oracle import pipeline, retrieval
    for q in range(queries_n):
        ref = results[q]
        retrieval.assert_ranking_matches(
            got["dense_rows"][q, :TOPK], got["dense_scores"][q, :TOPK], ref["dense_ids"],
            ref["dense_scores"], all_scores=ref.get("dense_all"), rtol=1e-5, atol=1e-6,
            what=f"bench dense q{q}")
        b_all = ref["bm25_all"]
        g_b = got["bm25_ids"][q, :TOPK]
        s_g, s_w = b_all[g_b], b_all[ref["bm25_docs"]]
        assert np.all(np.abs(s_g - s_w) <= 1e-5 * np.abs(s_w) + 1e-6), f"bench bm25 q{q}"
        c = int(got["counts"][q])
        pipeline.check_fused(got["ids"][q, :c], got["scores"][q, :c], got["dense_rows"][q, :TOPK],
                             g_b, (W_DENSE, W_BM25), WRRF_K, TOPK)
    return queries_n


def cpu_threads() -> int:
    """Threads the CPU port really uses: the BLAS pool behind np.dot (the CSR BM25 statement and the
    Python RRF loop are single-threaded, as in the reference)."""
    try:
        from threadpoolctl import threadpool_info
        n = max((int(p.get("num_threads", 0)) for p in threadpool_info()
                 if p.get("user_api") == "blas"), default=0)
        if n > 0:
            return n
    except Exception:
        pass
    return os.cpu_count() or 1


def workload_name(args) -> str:
    return (f"{args.chunks} chunks x {D}-d fp32 + BM25 {args.chunks} docs V={VOCAB} Zipf {ZIPF_S}, "
            f"{N_TERMS}-term queries, top-{TOPK} WRRF (w 5:1, k=40)")


# ---------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU port on this box's host cores; rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import torch
    synth = importlib.import_module("a-nice-rag_b200.synth")
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    emb, post = make_shard(torch, device, 0, args.chunks, 0)   # torch only generates the data
    emb_host = emb.cpu().numpy()
    post_host = {k: v.cpu().numpy() for k, v in post.items()}
    del emb, post
    idf = synth.idf_from_counts(args.chunks, post_host["nd"], EPS)
    avgdl = float(post_host["doc_len"].astype(np.int64).sum() / args.chunks)
    queries, terms, _ = make_queries(args.batch)
    emb_h, index = cpu_port_setup(emb_host, post_host, idf, avgdl)
    per_step = max(1, min(args.cpu_queries, args.batch) // 2)
    for _ in range(args.warmup):
        cpu_port_run(emb_h, index, queries, terms, 1)
    total = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_port_run(emb_h, index, queries, terms, per_step)
        total += dt
    qps = per_step * args.steps / total
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the CUDA arm's line; a step here is a bounded sample of its batch
        "config": {"workload": workload_name(args), "batch": args.batch,
                   "sample_queries_per_step": per_step},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} queries per step over the full corpus; numpy "
                                   "BLAS dot + CSR BM25 + Python RRF on a pre-stacked matrix"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pkg = importlib.import_module("a-nice-rag_b200")
    engine, native, synth = pkg.engine, pkg.native, importlib.import_module("a-nice-rag_b200.synth")
    sharded = importlib.import_module("a-nice-rag_b200.sharded")
    ctx = engine.context(local_rank)

    # ---- corpus shard ------------------------------------------------------------------
    lo, hi = sharded.shard_range(args.chunks, rank, world)
    emb, post = make_shard(torch, device, lo, hi, rank)
    nd, n_docs_total, avgdl = sharded.global_bm25_stats(
        post["nd"], int(post["doc_len"].to(torch.int64).sum()), hi - lo)
    idf = synth.idf_from_counts(n_docs_total, nd.cpu().numpy(), EPS)
    dense = engine.DenseIndex(emb, borrow=True)
    dense.set_shadow(args.dense_operands == "bf16")
    bm25 = engine.Bm25Index(post["term_ptr"], post["post_doc"], post["post_tf"], post["doc_len"],
                            idf, K1, B_PARAM, avgdl, n_terms=VOCAB, n_docs=hi - lo)
    n_postings = bm25.n_postings

    # ---- queries -------------------------------------------------------------------------
    B = args.batch
    q_host, t_host, off_host = make_queries(B)
    q_pin = torch.from_numpy(q_host).pin_memory()
    t_pin = torch.from_numpy(t_host.reshape(-1).copy()).pin_memory()
    off_pin = torch.from_numpy(off_host).pin_memory()
    q_dev, t_dev, off_dev = q_pin.to(device), t_pin.to(device), off_pin.to(device)
    out_ids = torch.empty((B, TOPK), dtype=torch.int32, device=device)
    out_scores = torch.empty((B, TOPK), dtype=torch.float64, device=device)
    out_counts = torch.empty((B,), dtype=torch.int32, device=device)
    h_ids = torch.empty((B, TOPK), dtype=torch.int32).pin_memory()
    h_scores = torch.empty((B, TOPK), dtype=torch.float64).pin_memory()
    h_counts = torch.empty((B,), dtype=torch.int32).pin_memory()
    shard = sharded.ShardedHybrid(dense, bm25, row_base=lo) if world > 1 else None

    def step_device(b=B):
        stream = engine.torch_stream_ptr()
        if world > 1:
            return shard.search(q_dev[:b], t_dev, off_dev, b, TOPK, W_DENSE, W_BM25, WRRF_K, TOPK)
        native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle, q_dev.data_ptr(),
                    t_dev.data_ptr(), off_dev.data_ptr(), b, TOPK, TOPK, None, None, None, 0,
                    W_DENSE, W_BM25, WRRF_K, TOPK, out_ids.data_ptr(), out_scores.data_ptr(),
                    out_counts.data_ptr(), None, None, None, None, stream)
        return out_ids, out_scores, out_counts

    def step_e2e(b=B):
        """Host buffers in, host buffers out (the call synchronises before returning)."""
        stream = engine.torch_stream_ptr()
        if world > 1:
            qd = q_pin[:b].to(device, non_blocking=True)
            td = t_pin.to(device, non_blocking=True)
            od = off_pin.to(device, non_blocking=True)
            ids, scores, counts = shard.search(qd, td, od, b, TOPK, W_DENSE, W_BM25, WRRF_K, TOPK)
            h_ids[:b].copy_(ids, non_blocking=True)
            h_scores[:b].copy_(scores, non_blocking=True)
            h_counts[:b].copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return
        native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle, q_pin.data_ptr(),
                    t_pin.data_ptr(), off_pin.data_ptr(), b, TOPK, TOPK, None, None, None, 0,
                    W_DENSE, W_BM25, WRRF_K, TOPK, h_ids.data_ptr(), h_scores.data_ptr(),
                    h_counts.data_ptr(), None, None, None, None, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """ms for `steps` calls: CUDA events on the launching stream, max over ranks."""
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- warm-up, then the timed regions --------------------------------------------------
    # clocks are sampled from the warm-up to the end of the latency loop (the timed regions
    # alone can be shorter than one nvidia-smi sampling period)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    native.call("anr_ctx_profile_enable", ctx.handle, 1)
    for kind in (0, 1, 2):
        native.call("anr_ctx_profile_read", ctx.handle, kind, None, None)   # reset counters
    ms_dev = timed(step_device, args.steps)
    scan_ms, scan_n, bm_ms, bm_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    pass_ms, pass_n = C.c_double(), C.c_int64()
    native.call("anr_ctx_profile_read", ctx.handle, 0, C.byref(scan_ms), C.byref(scan_n))
    native.call("anr_ctx_profile_read", ctx.handle, 1, C.byref(bm_ms), C.byref(bm_n))
    native.call("anr_ctx_profile_read", ctx.handle, 2, C.byref(pass_ms), C.byref(pass_n))
    native.call("anr_ctx_profile_enable", ctx.handle, 0)
    ms_e2e = timed(step_e2e, args.steps)

    # ---- filtered variant: a "CG,NG"-style source filter keeping ~2/3 of the rows,
    #      passed as row / document bit masks (search_engine.py:36-55, :221-231) ----------------
    ms_filtered = None
    if world == 1:
        rng = np.random.default_rng(99)
        keep = rng.random(hi - lo) < (2.0 / 3.0)
        words = torch.from_numpy(engine.pack_mask(keep).view(np.int32)).to(device)

        def step_filtered():
            native.call("anr_hybrid_search", ctx.handle, dense.handle, bm25.handle,
                        q_dev.data_ptr(), t_dev.data_ptr(), off_dev.data_ptr(), B, TOPK, TOPK,
                        words.data_ptr(), words.data_ptr(), None, 0, W_DENSE, W_BM25, WRRF_K, TOPK,
                        out_ids.data_ptr(), out_scores.data_ptr(), out_counts.data_ptr(), None, None,
                        None, None, engine.torch_stream_ptr())
        for _ in range(3):
            step_filtered()
        ms_filtered = timed(step_filtered, args.steps)

    # ---- batch-1 latency through the C-ABI with host buffers (p50 of wall-clock per call) -----
    lat = []
    if world == 1:
        for i in range(args.latency_iters + 20):
            t0 = time.perf_counter()
            step_e2e(1)
            if i >= 20:
                lat.append(1e3 * (time.perf_counter() - t0))
    ms_b1_dev = timed(lambda: step_device(1), max(args.steps, 50)) / max(args.steps, 50)
    # keep the GPU under the same load until at least a few clock samples exist
    # (a fixed count derived from the max-reduced step time: every rank runs the same number
    # of collectives)
    for _ in range(int(min(2000, max(10, 1000.0 / max(ms_dev / args.steps, 1e-3))))):
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    # batch-1 again through the fp32 CUDA-core scan (north_star kernel (1)): without the bf16 shadow
    ms_b1_scan, lat_scan = None, []
    if args.dense_operands == "bf16":
        dense.set_shadow(False)
        for _ in range(5):
            step_device(1)
        ms_b1_scan = timed(lambda: step_device(1), max(args.steps, 50)) / max(args.steps, 50)
        if world == 1:
            for i in range(args.latency_iters + 20):
                t0 = time.perf_counter()
                step_e2e(1)
                if i >= 20:
                    lat_scan.append(1e3 * (time.perf_counter() - t0))
        dense.set_shadow(True)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_kind = load_peaks()
    traffic = None        # dram bytes per launch, from the committed ncu captures
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    rows_local = hi - lo
    # anr_dense_gemm.cu takes batches > 32, and every batch when a bf16 shadow exists
    gemm = (B > 32 or args.dense_operands == "bf16") and rows_local >= 80_000
    shadow = gemm and args.dense_operands == "bf16"
    tile_q = 64 if B <= 64 else (128 if B <= 128 else 256)
    dense_kernel = (f"dense_gemm_kernel<{tile_q}, {'bf16' if shadow else 'tf32'}>" if gemm else
                    (("dense_tc_pair_kernel" if B > 32 else "dense_tc_kernel<0>") if B > 8
                     else "dense_scan_kernel<1, 4, 0>"))
    # algorithmic bytes of one launch: every row once, in the operand type the kernel reads
    scan_bytes = rows_local * D * (2 if shadow else 4)
    scan_avg_ms = scan_ms.value / max(scan_n.value, 1)
    achieved = scan_bytes / (scan_avg_ms * 1e-3) / 1e9 if scan_avg_ms > 0 else 0.0
    bm_avg_ms = bm_ms.value / max(bm_n.value, 1)
    # BM25: 8 B (doc id + weight) per posting of every query-term occurrence of the batch
    nd_local = post["nd"].cpu().numpy()
    bm_bytes = 8 * int(nd_local[t_host.reshape(-1)].astype(np.int64).sum())
    bm_achieved = bm_bytes / (bm_avg_ms * 1e-3) / 1e9 if bm_avg_ms > 0 else 0.0
    if os.path.exists(tpath) and args.chunks == 1_000_000 and world == 1:
        with open(tpath) as fh:
            tj = json.load(fh)["bytes_per_launch"]
        traffic = tj.get(dense_kernel)
        bm_traffic = tj.get("bm25_score_kernel<0> batch 64") if B == 64 else None
    else:
        bm_traffic = None
    dense_roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                  "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                  "kernel": dense_kernel, "bytes_per_launch": scan_bytes,
                  "avg_launch_ms": scan_avg_ms, "launches": int(scan_n.value),
                  "share_of_step": scan_ms.value / ms_dev if ms_dev else None}
    bm_roof = {"bound": "hbm", "achieved": bm_achieved, "peak": peak, "unit": "GB/s",
               "frac": bm_achieved / peak, "traffic": bm_traffic, "peak_source": peak_kind,
               "kernel": "bm25_score_kernel", "bytes_per_launch": bm_bytes,
               "avg_launch_ms": bm_avg_ms, "launches": int(bm_n.value),
               "share_of_step": bm_ms.value / ms_dev if ms_dev else None}
    # BM25 runs on the library's side stream UNDER the dense pass: its in-step event time includes
    # waiting for SMs the dense kernel holds, so it is also timed alone (same batch, same kernels)
    bm_alone_ms = None
    if world == 1:
        k_sc = torch.empty((B, TOPK), dtype=torch.float32, device=device)
        k_id = torch.empty((B, TOPK), dtype=torch.int32, device=device)
        k_ct = torch.empty((B,), dtype=torch.int32, device=device)

        def bm25_only():
            native.call("anr_bm25_search", ctx.handle, bm25.handle, t_dev.data_ptr(),
                        off_dev.data_ptr(), B, TOPK, None, None, 0, k_sc.data_ptr(), k_id.data_ptr(),
                        k_ct.data_ptr(), engine.torch_stream_ptr())
        native.call("anr_ctx_profile_enable", ctx.handle, 1)
        for _ in range(3):
            bm25_only()
        native.call("anr_ctx_profile_read", ctx.handle, 1, None, None)
        for _ in range(10):
            bm25_only()
        torch.cuda.synchronize()
        a_ms, a_n = C.c_double(), C.c_int64()
        native.call("anr_ctx_profile_read", ctx.handle, 1, C.byref(a_ms), C.byref(a_n))
        native.call("anr_ctx_profile_enable", ctx.handle, 0)
        bm_alone_ms = a_ms.value / max(a_n.value, 1)
        # the roofline figures of the BM25 launch come from this pass; its in-step event time
        # (which includes waiting for SMs the dense kernel holds) is kept beside them
        bm_roof["overlapped"] = ("runs on a side stream under the dense pass: avg_launch_ms / achieved "
                                 "/ frac are the same launch timed right after the timed region "
                                 "without the dense pass; in_step_ms is its event time inside the step")
        bm_roof["in_step_ms"] = bm_avg_ms
        bm_roof["in_step_share"] = bm_roof["share_of_step"]
        if bm_alone_ms:
            bm_roof["share_of_step"] = bm_alone_ms * max(bm_n.value, 1) / ms_dev if ms_dev else None
            bm_roof["avg_launch_ms"] = bm_alone_ms
            bm_roof["achieved"] = bm_bytes / (bm_alone_ms * 1e-3) / 1e9
            bm_roof["frac"] = bm_roof["achieved"] / peak
    # Co-resident configuration (ANR_GEMM_BESIDE_STAGES / ANR_GEMM_MAX_STAGES: the dense main kernel
    # shares every SM with the BM25 scan): its in-step time includes that sharing,
    # so the same batch is also timed through a dense-only call, as is done for BM25 above.  Only
    # active when one of the two knobs is set: the default run does not execute this block.
    if world == 1 and (os.environ.get("ANR_GEMM_BESIDE_STAGES") or os.environ.get("ANR_GEMM_MAX_STAGES")):
        try:
            d_sc = torch.empty((B, TOPK), dtype=torch.float32, device=device)
            d_id = torch.empty((B, TOPK), dtype=torch.int32, device=device)
            d_ct = torch.empty((B,), dtype=torch.int32, device=device)

            def dense_only():
                native.call("anr_dense_search", ctx.handle, dense.handle, q_dev.data_ptr(), B, TOPK,
                            None, 0, d_sc.data_ptr(), d_id.data_ptr(), d_ct.data_ptr(),
                            engine.torch_stream_ptr())
            native.call("anr_ctx_profile_enable", ctx.handle, 1)
            for _ in range(3):
                dense_only()
            native.call("anr_ctx_profile_read", ctx.handle, 0, None, None)
            for _ in range(10):
                dense_only()
            torch.cuda.synchronize()
            a_ms, a_n = C.c_double(), C.c_int64()
            native.call("anr_ctx_profile_read", ctx.handle, 0, C.byref(a_ms), C.byref(a_n))
            native.call("anr_ctx_profile_enable", ctx.handle, 0)
            alone = a_ms.value / max(a_n.value, 1)
            dense_roof["co_resident"] = ("shares every SM with the BM25 scan inside the step: "
                                         "avg_launch_ms / achieved / frac are in-step; alone_* is the "
                                         "same batch through a dense-only call")
            dense_roof["alone_ms"] = alone
            dense_roof["alone_achieved"] = scan_bytes / (alone * 1e-3) / 1e9 if alone > 0 else None
            dense_roof["alone_frac"] = dense_roof["alone_achieved"] / peak if alone > 0 else None
        except Exception as exc:
            dense_roof["alone_error"] = repr(exc)[:200]
    # dominant kernel = larger exclusive time (BM25 judged by its time alone when overlapped)
    bm_cmp = bm_alone_ms * max(bm_n.value, 1) if bm_alone_ms else bm_ms.value
    dominant = dense_roof if scan_ms.value >= bm_cmp else bm_roof

    # ---- the same step replayed from a CUDA graph (graph.HybridGraph) -----------------
    # Reported beside the eager numbers, which stay the headline: one launch instead of ~20 plus
    # the fork/join event traffic.  Results must equal the eager call's bit for bit.
    graph_rec = None
    if world == 1:
        try:
            graph_mod = importlib.import_module("a-nice-rag_b200.graph")
            graph_rec = {}
            n_it = max(args.steps, 50)
            for b_g in sorted({1, B}):
                hg = graph_mod.HybridGraph(dense, bm25, b_g, N_TERMS * b_g, TOPK, TOPK, W_DENSE,
                                           W_BM25, WRRF_K, TOPK)
                hg.load(q_dev[:b_g], t_dev[:N_TERMS * b_g], off_dev[:b_g + 1])
                for _ in range(5):
                    hg.replay()
                step_device(b_g)
                torch.cuda.synchronize()
                same = bool(torch.equal(hg.ids, out_ids[:b_g]) and
                            torch.equal(hg.scores, out_scores[:b_g]) and
                            torch.equal(hg.counts, out_counts[:b_g]))
                ms_g = timed(hg.replay, n_it) / n_it
                graph_rec[f"batch{b_g}"] = {"replay_ms": ms_g, "queries_per_s": 1e3 * b_g / ms_g,
                                            "identical_to_eager": same}
                del hg
        except Exception as exc:   # never lose the headline line to the optional replay leg
            graph_rec = {"error": repr(exc)[:300]}

    # ---- OPT-IN (ANR_BENCH_PIPELINE=1): host buffers in / out with two batches in flight
    #      (graph.HybridPipeline, experimental until it has run on a B200) --------------------------
    pipe_rec = None
    if world == 1 and os.environ.get("ANR_BENCH_PIPELINE") == "1":
        try:
            graph_mod = importlib.import_module("a-nice-rag_b200.graph")
            pipe = graph_mod.HybridPipeline(dense, bm25, B, N_TERMS * B, TOPK, TOPK, W_DENSE, W_BM25,
                                            WRRF_K, TOPK, depth=2)
            t_flat = t_host.reshape(-1)

            def run_pipe(n_steps):
                for i in range(n_steps):
                    if i >= 2:
                        pipe.collect()
                    pipe.submit(q_host, t_flat, off_host)
                last = None
                for _ in range(min(2, n_steps)):
                    last = pipe.collect()
                return last
            run_pipe(6)
            step_device()
            torch.cuda.synchronize()
            last = run_pipe(3)
            same = bool(np.array_equal(last[0], out_ids.cpu().numpy()) and
                        np.array_equal(last[1], out_scores.cpu().numpy()))
            n_it = max(args.steps, 50)
            t0 = time.perf_counter()
            run_pipe(n_it)
            dt = time.perf_counter() - t0
            pipe_rec = {"value": B * n_it / dt, "unit": "queries/s", "ms_per_step": 1e3 * dt / n_it,
                        "depth": 2, "identical_to_eager": same,
                        "timing": "host wall clock over the whole loop (every result read on the host)"}
        except Exception as exc:
            pipe_rec = {"error": repr(exc)[:300]}

    # ---- CPU baseline + parity of this very batch (N = 1) --------------------------------------
    cpu = None
    checked = 0
    if world == 1 and not args.no_cpu_baseline:
        emb_host = emb.cpu().numpy()
        post_host = {k: v.cpu().numpy() for k, v in post.items()}
        emb_h, index = cpu_port_setup(emb_host, post_host, idf, avgdl)
        nq_cpu = min(args.cpu_queries, B)
        cpu_port_run(emb_h, index, q_host, t_host, 1)                       # warm caches / BLAS
        dt, results = cpu_port_run(emb_h, index, q_host, t_host, nq_cpu)
        cpu = {"value": nq_cpu / dt, "unit": "queries/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"{nq_cpu} of the batch's {B} queries over the full {args.chunks}-chunk "
                         "corpus; numpy BLAS dot + CSR BM25 + Python RRF on a pre-stacked matrix "
                         "(the reference's per-query np.stack and pandas work NOT included)"}
        # the WHOLE batch through the same kernels as the timed steps; the sampled queries are checked
        got = engine.hybrid_search(dense, bm25, q_host, [list(map(int, t)) for t in t_host],
                                   TOPK, TOPK, W_DENSE, W_BM25, WRRF_K, TOPK, want_lists=True)
        for r in results:   # full dense score vector only where ids differ is costly: recompute lazily
            r["dense_all"] = None
        for q in range(nq_cpu):
            if not np.array_equal(got["dense_rows"][q], results[q]["dense_ids"]):
                results[q]["dense_all"] = emb_h @ q_host[q]
        checked = check_against_cpu(results, got, nq_cpu)

    line = {
        "metric": METRIC, "value": B * args.steps / (ms_dev * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.chunks_per_gpu else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "batch": B, "sharding": f"chunks/{world}",
                   "l2": "inputs larger than L2 (corpus shard >= 0.5 GB per GPU)",
                   "bm25_postings_local": n_postings},
        "e2e": {"value": B * args.steps / (ms_e2e * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": int(q_pin.numel() * 4 + t_pin.numel() * 4 + off_pin.numel() * 4),
                "d2h_bytes_per_step": int(h_ids.numel() * 4 + h_scores.numel() * 8 + h_counts.numel() * 4)},
        # kernels of this repo launched per step: dense (CUDA-core scan + final | tcgen05:
        # pre-pass(es) + threshold + scan + rescore + flag compaction + flagged rescan + its merge),
        # BM25 score + final, weights, WRRF; sharded: + the two merges of anr_sharded_fuse
        # GEMM path: [query -> bf16] + sample pass + thresholds + GEMM + rescore + flag
        # compaction + flagged rescan + its merge; BM25: [sample launch] + scan + final top-k;
        # then the weight upload kernel and WRRF
        "gpu_launches": int(args.steps * (
            (7 + (1 if shadow else 0) if gemm else ((8 if B > 32 else 7) if B > 8 else 2))
            + (3 if B >= 16 and (rows_local + 6143) // 6144 >= 32 else 2) + 2
            + (2 if world > 1 else 0))),
        # the kernel with the largest share of the step; the other one follows
        "roofline": dominant,
        "roofline_other": bm_roof if dominant is dense_roof else dense_roof,
        "dense_operands": ("bf16 shadow copy" if shadow else "fp32 words as tf32") if (B > 8 or gemm) else "fp32",
        "dense_tc_pass": {"what": "sample pre-pass + threshold + scan + exact rescoring",
                          "avg_ms": pass_ms.value / max(pass_n.value, 1), "passes": int(pass_n.value),
                          "share_of_step": pass_ms.value / ms_dev if ms_dev else None},
        "batch1": {"device_ms": ms_b1_dev, "device_qps": 1e3 / ms_b1_dev if ms_b1_dev else None,
                   "e2e_p50_ms": statistics.median(lat) if lat else None,
                   "e2e_p95_ms": (sorted(lat)[int(0.95 * len(lat))] if lat else None),
                   "dense_path": ("bf16 shadow GEMM pass + exact fp32 rescoring" if shadow
                                  else "fp32 CUDA-core scan")},
        "batch1_fp32_scan": ({"device_ms": ms_b1_scan,
                              "e2e_p50_ms": statistics.median(lat_scan) if lat_scan else None,
                              "hbm_gbs": rows_local * D * 4 / (ms_b1_scan * 1e-3) / 1e9,
                              "frac_of_peak": rows_local * D * 4 / (ms_b1_scan * 1e-3) / 1e9 / peak}
                             if ms_b1_scan else None),
        "filtered": ({"what": "same step with row / document masks keeping 2/3 of the corpus "
                              "(the reference's source-prefix filter)",
                      "value": B * args.steps / (ms_filtered * 1e-3), "unit": "queries/s",
                      "ms_per_step": ms_filtered / args.steps} if ms_filtered else None),
        "cuda_graph": graph_rec,
        "e2e_pipelined": pipe_rec,
        "clocks": clocks, "parity_checked_queries": checked,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global VOCAB
    args = parse_args()
    VOCAB = args.vocab
    if args.chunks_per_gpu:
        args.chunks = args.chunks_per_gpu * dist_env()[2]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
