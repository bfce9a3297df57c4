#!/bin/bash
# Round 2, GPU call 9: BM25 as ONE folded launch behind the dense pre-pass (no separate sample
# launch holding the SMs' shared memory while the dense sample pass wants it) x ring depth.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --cpu-queries 8 > gpurun_out/c9_${name}.json 2> gpurun_out/c9_${name}.err
  echo "$name rc=$?"; tail -2 gpurun_out/c9_${name}.err
}
run fold_ring4 ANR_BM25_FOLD=1
run fold_ring3 ANR_BM25_FOLD=1 ANR_GEMM_BESIDE_STAGES=3
run fold_ring3_div8 ANR_BM25_FOLD=1 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_HEAD_DIV=8
run fold_ring5 ANR_BM25_FOLD=1 ANR_GEMM_BESIDE_STAGES=5
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c9_*.json")):
    try:
        d = json.load(open(f))
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "parity", d["parity_checked_queries"], d.get("parity_error"),
              "b1", round(d["batch1"]["device_ms"], 4), "2inflight", (p.get("two_in_flight") or {}).get("ms_per_step"), "pipe", (p.get("e2e_pipelined") or {}).get("value"))
        print("   timeline", {k: v for k, v in (d.get("timeline") or {}).items() if k != "unit"})
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
