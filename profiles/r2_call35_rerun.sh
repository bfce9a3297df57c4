#!/bin/bash
# Round 2, GPU call 35: anr_ctx_last_rerun (suite incl. its test) + headline with `reruns`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c35_suite.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/c35_suite.log
timeout 300 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 10 --legs headline --no-cpu-baseline \
    > gpurun_out/c35_bench.json 2> gpurun_out/c35_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/c35_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c35_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), round(d["ms_per_step"], 4), "reruns", d.get("reruns"), "parity", d.get("parity_error"))
PY
