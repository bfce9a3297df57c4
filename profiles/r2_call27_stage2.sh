#!/bin/bash
# Round 2, GPU call 27: stage-2 variants (segment cache + adaptive refill): headline + BM25-only leg.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-a}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bm25" > gpurun_out/c27_${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c27_${TAG}_tests.log
timeout 400 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline,big --no-cpu-baseline \
    > gpurun_out/c27_${TAG}.json 2> gpurun_out/c27_${TAG}.err
echo "bench rc=$?"; tail -2 gpurun_out/c27_${TAG}.err
python - $TAG <<'PY'
import json, sys
d = json.loads([l for l in open(f"gpurun_out/c27_{sys.argv[1]}.json") if l.startswith("{")][-1])
p = d.get("pipelined") or {}
o = d["roofline_other"]
print("value", round(d["value"]), round(d["ms_per_step"], 4), "bm25 alone", round(o["alone_ms"], 4), "in-step", round(o["in_step_ms"], 4),
      "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
l = d["legs"]
print("  config4", round(l["config4"]["ms_per_batch"], 3), "ms  parity", l["config4"].get("parity_checked_queries"), "| config3 bf16", round(l["config3"]["bf16"]["ms_per_batch"], 2),
      "| b1_10M", round(l["batch1_10M"]["bf16_shadow"]["device_ms"], 3), l.get("parity_error"))
PY
