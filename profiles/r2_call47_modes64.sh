#!/bin/bash
# Round 2, GPU call 47: the 64-query pass with the query slab shared by a CTA pair (ANR_GEMM_MODE=1:
# multicast halves, =2: cta_group::2 MMAs) against one CTA per tile (=0): the L2 delivers the query
# slab again for every corpus slab (+50 % L2 traffic at 64 bf16 queries).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for M in 0 1 2; do
ANR_GEMM_MODE=$M timeout 300 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --cpu-queries 8 \
    > gpurun_out/c47_bench_m$M.json 2> gpurun_out/c47_bench_m$M.err
echo "bench mode=$M rc=$?"
python - $M <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/c47_bench_m%s.json" % sys.argv[1]) if l.startswith("{")][-1])
p = d.get("pipelined") or {}
r = d["roofline"]
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_checked_queries"), d.get("parity_error"))
print("  dense kernel in step", r.get("avg_launch_ms"), "alone", r.get("alone_ms"), "frac", r.get("frac"), r.get("alone_frac"))
print("  batch1", d["batch1"]["device_ms"], "clocks", d["clocks"], "reruns", d.get("reruns"))
PY
done
