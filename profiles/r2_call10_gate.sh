#!/bin/bash
# Round 2, GPU call 10: gated hybrid schedule (BM25 held until the dense main kernel is resident).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10_suite.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/c10_suite.log
run() {
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --cpu-queries 8 > gpurun_out/c10_${name}.json 2> gpurun_out/c10_${name}.err
  echo "$name rc=$?"; tail -2 gpurun_out/c10_${name}.err
}
run gate_ring4 A=1
run gate_ring3 ANR_GEMM_BESIDE_STAGES=3
run gate_ring5 ANR_GEMM_BESIDE_STAGES=5
run nogate_ring4 ANR_HYBRID_GATE=0
run gate_ring3_div8 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_HEAD_DIV=8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c10_*.json")):
    try:
        d = json.load(open(f))
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "parity", d["parity_checked_queries"], d.get("parity_error"),
              "b1", round(d["batch1"]["device_ms"], 4), "e2e", round(d["e2e"]["value"]), "graph", round((d.get("cuda_graph") or {}).get("batch64", {}).get("replay_ms", 0), 4),
              "2inflight", (p.get("two_in_flight") or {}).get("ms_per_step"), "pipe", (p.get("e2e_pipelined") or {}).get("value"), p.get("error"))
        print("   timeline", {k: v for k, v in (d.get("timeline") or {}).items() if k != "unit"})
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
