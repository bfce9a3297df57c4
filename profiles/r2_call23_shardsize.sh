#!/bin/bash
# Round 2, GPU call 23: one GPU holding what ONE of 8 / 4 / 2 ranks holds of the 1M corpus (125k /
# 250k / 500k chunks): the fixed costs of a sharded step without paying for 8 GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 125000 250000 500000; do
  timeout 200 python bench.py --chunks $n --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --no-cpu-baseline > gpurun_out/c23_${n}.json 2> gpurun_out/c23_${n}.err
  echo "$n rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c23_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "bm25 alone", round(d["roofline_other"].get("alone_ms") or d["roofline"].get("alone_ms") or 0, 4),
              "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4))
        print("   roofline", d["roofline"]["kernel"], round(d["roofline"]["avg_launch_ms"], 4), "| other", d["roofline_other"]["kernel"], d["roofline_other"].get("in_step_ms"), d["roofline_other"].get("avg_launch_ms"))
        print("   timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
    except Exception as e:
        print(f, "ERR", e)
PY
