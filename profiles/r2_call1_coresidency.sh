#!/bin/bash
# Round 2, GPU call 1: (a) the whole GPU suite with the co-resident configuration
# (ANR_GEMM_BESIDE_STAGES=4) and with the persistent BM25 grid, (b) bench variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ANR_GEMM_BESIDE_STAGES=4 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c1_suite_beside4.log 2>&1
echo "suite beside4 rc=$?"; tail -3 gpurun_out/c1_suite_beside4.log
ANR_GEMM_BESIDE_STAGES=4 ANR_BM25_PERSISTENT=3 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c1_suite_persist3.log 2>&1
echo "suite beside4+persist3 rc=$?"; tail -3 gpurun_out/c1_suite_persist3.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --latency-iters 20 --no-cpu-baseline \
    > gpurun_out/c1_${name}.json 2> gpurun_out/c1_${name}.err
  echo "$name rc=$?"
}
run default A=1
run beside4 ANR_GEMM_BESIDE_STAGES=4
run beside3 ANR_GEMM_BESIDE_STAGES=3
run persist6 ANR_BM25_PERSISTENT=6
run beside4_persist3 ANR_GEMM_BESIDE_STAGES=4 ANR_BM25_PERSISTENT=3
run beside3_persist4 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_PERSISTENT=4
run beside4_persist3_tile4096 ANR_GEMM_BESIDE_STAGES=4 ANR_BM25_PERSISTENT=4 ANR_BM25_TILE=4096
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c1_*.json")):
    try:
        d = json.load(open(f))
        ro, rd = d["roofline_other"], d["roofline"]
        if rd["kernel"].startswith("bm25"): ro, rd = rd, ro
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "dense", round(rd["avg_launch_ms"], 4),
              "alone", rd.get("alone_ms"), "bm25 in-step", round(ro.get("in_step_ms") or 0, 4),
              "bm25 alone", round(ro["avg_launch_ms"], 4), "b1", round(d["batch1"]["device_ms"], 4),
              "e2e", round(d["e2e"]["value"]), (d.get("cuda_graph") or {}).get("batch64"))
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
