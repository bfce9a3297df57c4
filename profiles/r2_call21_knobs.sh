#!/bin/bash
# Round 2, GPU call 21: head-row threshold (more terms as dense rows = one-load lookups) under the
# candidate-driven BM25 path; headline only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --no-cpu-baseline > gpurun_out/c21_${name}.json 2> gpurun_out/c21_${name}.err
  echo "$name rc=$?"
}
run div4 A=1
run div8 ANR_BM25_HEAD_DIV=8
run div16 ANR_BM25_HEAD_DIV=16
run div32 ANR_BM25_HEAD_DIV=32
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c21_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "bm25 alone", round(d["roofline_other"]["alone_ms"], 4), "in-step", round(d["roofline_other"]["in_step_ms"], 4),
              "dense in-step", round(d["roofline"]["avg_launch_ms"], 4), "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
