"""Times the dense scan kernel alone (CUDA events via the library's profiling hook) for
nq = 1, 2, 4, 8 over an n x 1024 corpus; the command ncu is pointed at for the scan kernel."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("a-nice-rag_b200")
engine, native, synth = pkg.engine, pkg.native, importlib.import_module("a-nice-rag_b200.synth")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nqs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
d, k = 1024, 10
dev = torch.device("cuda", 0)
emb = synth.unit_vectors_torch(n, d, 1234, dev)
index = engine.DenseIndex(emb, borrow=True)
ctx = engine.context(0)
native.call("anr_ctx_profile_enable", ctx.handle, 1)
for nq in nqs:
    q = torch.from_numpy(synth.unit_vectors(nq, d, seed=4321)).to(dev)
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
    counts = torch.empty((nq,), dtype=torch.int32, device=dev)
    for it in range(iters + 3):
        if it == 3:
            native.call("anr_ctx_profile_read", ctx.handle, 0, None, None)
            native.call("anr_ctx_profile_read", ctx.handle, 2, None, None)
        native.call("anr_dense_search", ctx.handle, index.handle, q.data_ptr(), nq, k, None, 0,
                    scores.data_ptr(), rows.data_ptr(), counts.data_ptr(), None)
    ms, cnt = C.c_double(), C.c_int64()
    native.call("anr_ctx_profile_read", ctx.handle, 0, C.byref(ms), C.byref(cnt))
    avg = ms.value / max(cnt.value, 1)
    pms, pcnt = C.c_double(), C.c_int64()
    native.call("anr_ctx_profile_read", ctx.handle, 2, C.byref(pms), C.byref(pcnt))
    extra = f"; whole tensor-core pass {pms.value / pcnt.value:.4f} ms" if pcnt.value else ""
    print(f"nq={nq} rw={os.environ.get('ANR_SCAN_RW', 'auto')} scan kernel avg {avg:.4f} ms over {cnt.value} launches "
          f"-> {n * d * 4 / avg / 1e6:.1f} GB/s{extra}", flush=True)
