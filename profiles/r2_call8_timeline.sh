#!/bin/bash
# Round 2, GPU call 8: the new cross-stream test + graph tests, then the headline with the step
# timeline for ring 3 / ring 4 (full sample), full parity.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_zz_pipeline.py -x -q > gpurun_out/c8_suite.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/c8_suite.log
run() {
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    > gpurun_out/c8_${name}.json 2> gpurun_out/c8_${name}.err
  echo "$name rc=$?"; tail -2 gpurun_out/c8_${name}.err
}
run ring4 A=1
run ring3 ANR_GEMM_BESIDE_STAGES=3
run ring0 ANR_GEMM_BESIDE_STAGES=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c8_*.json")):
    try:
        d = json.load(open(f))
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "parity", d["parity_checked_queries"], d.get("parity_error"),
              "2inflight", (p.get("two_in_flight") or {}).get("ms_per_step"), "pipe", (p.get("e2e_pipelined") or {}).get("value"))
        print("   timeline", d.get("timeline"))
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
