#!/bin/bash
# Round 2, GPU call 5: BM25 light sample launch x dense ring depth (4 BM25 CTAs per SM beside a
# 3-stage ring now that the GEMM epilogue staging is sized by its warp count); BM25 suite first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bm25 or hybrid or sharded or graph or pipeline or dropin or retrieve" > gpurun_out/c5_suite.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/c5_suite.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 150 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 \
    --no-cpu-baseline --legs headline \
    > gpurun_out/c5_${name}.json 2> gpurun_out/c5_${name}.err
  echo "$name rc=$?"
}
run ring4_light A=1
run ring4_full ANR_BM25_LIGHT_SAMPLE=0
run ring3_light ANR_GEMM_BESIDE_STAGES=3
run ring3_full ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_LIGHT_SAMPLE=0
run ring3_light_tile5632 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_TILE=5632
run ring3_light_div3 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_HEAD_DIV=3
run ring3_light_div2 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_HEAD_DIV=2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c5_*.json")):
    try:
        d = json.load(open(f))
        ro, rd = d["roofline_other"], d["roofline"]
        if rd["kernel"].startswith("bm25"): ro, rd = rd, ro
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "dense", round(rd["avg_launch_ms"], 4),
              "bm25 in-step", round(ro.get("in_step_ms") or 0, 4), "alone", round(ro.get("alone_ms") or 0, 4),
              "b1", round(d["batch1"]["device_ms"], 4), "e2e", round(d["e2e"]["value"]),
              "graph", round((d.get("cuda_graph") or {}).get("batch64", {}).get("replay_ms", 0), 4),
              "2inflight", (p.get("two_in_flight") or {}).get("ms_per_step"), (p.get("two_in_flight") or {}).get("identical_to_eager"),
              "pipe", (p.get("e2e_pipelined") or {}).get("value"), p.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python profiles/f4_evaluator_bench.py --queries 1024 > gpurun_out/c5_f4.json 2> gpurun_out/c5_f4.err
echo "f4 rc=$?"; cat gpurun_out/c5_f4.json; tail -3 gpurun_out/c5_f4.err
exit 0
