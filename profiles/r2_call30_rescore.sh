#!/bin/bash
# Round 2, GPU call 30: rescoring kernel with 1024 threads: dense tests + headline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_graph.py tests/test_gpu_full_size.py -x -q > gpurun_out/c30_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c30_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --no-cpu-baseline \
    > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c30_bench.json") if l.startswith("{")][-1])
p = d.get("pipelined") or {}
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
print("  batch1", d["batch1"])
PY
