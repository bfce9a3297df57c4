"""Correctness + timing probe of the tiled-GEMM dense path (anr_dense_gemm.cu).

Usage: python profiles/gemm_probe.py <n_rows> <nq> <k> [shadow 0|1] [iters] [check_queries]
Compares a sample of queries against torch fp32 (allow_tf32 off) on the same device tensor and
prints the main GEMM kernel time (library event hooks) and the whole-call time.
"""
import ctypes as C
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("a-nice-rag_b200")
engine, native, synth = pkg.engine, pkg.native, importlib.import_module("a-nice-rag_b200.synth")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
shadow = int(sys.argv[4]) if len(sys.argv) > 4 else 0
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
n_check = int(sys.argv[6]) if len(sys.argv) > 6 else 8
d = 1024
dev = torch.device("cuda", 0)
emb = synth.unit_vectors_torch(n, d, 1234, dev)
index = engine.DenseIndex(emb, borrow=True)
if shadow:
    index.set_shadow(True)
ctx = engine.context(0)
native.call("anr_ctx_profile_enable", ctx.handle, 1)
q = torch.from_numpy(synth.unit_vectors(nq, d, seed=4321)).to(dev)
scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
counts = torch.empty((nq,), dtype=torch.int32, device=dev)
stream = engine.torch_stream_ptr()


def run():
    native.call("anr_dense_search", ctx.handle, index.handle, q.data_ptr(), nq, k, None, 0,
                scores.data_ptr(), rows.data_ptr(), counts.data_ptr(), stream)


for _ in range(2):
    run()
torch.cuda.synchronize()
native.call("anr_ctx_profile_read", ctx.handle, 0, None, None)
native.call("anr_ctx_profile_read", ctx.handle, 2, None, None)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    run()
e.record()
torch.cuda.synchronize()
call_ms = s.elapsed_time(e) / iters
ms, cnt = C.c_double(), C.c_int64()
native.call("anr_ctx_profile_read", ctx.handle, 0, C.byref(ms), C.byref(cnt))
kern_ms = ms.value / max(cnt.value, 1)

torch.backends.cuda.matmul.allow_tf32 = False
bad = 0
step = max(nq // max(n_check, 1), 1)
checked = 0
for qi in range(0, nq, step):
    ref = torch.mv(emb, q[qi])
    top = torch.topk(ref, k)
    got = rows[qi].long()
    ok = int(counts[qi]) == k
    if ok and not bool((got == top.indices).all()):
        diff = (ref[got] - top.values).abs().max().item()
        ok = diff <= 1e-5 * top.values.abs().max().item() + 1e-6
    ok = ok and torch.allclose(scores[qi], ref[got], rtol=1e-5, atol=1e-6)
    bad += 0 if ok else 1
    checked += 1
print(json.dumps({
    "n": n, "nq": nq, "k": k, "operands": "bf16 shadow" if shadow else "tf32 on fp32",
    "gemm_kernel_ms": kern_ms, "kernel_launches": cnt.value, "call_ms": call_ms,
    "queries_per_s": nq / call_ms * 1e3,
    "kernel_tflops": 2.0 * nq * n * d / kern_ms / 1e9,
    "kernel_hbm_gbs_fp32_bytes": n * d * 4 / kern_ms / 1e6,
    "checked": checked, "mismatched": bad}), flush=True)
