#!/bin/bash
# Co-residency experiment, second take: capped GEMM ring WITH the max-shared carveout preference.
cd "$(dirname "$0")/.."
# ring 4 with the CPU parity check of 2 queries, ring 3 without
ANR_GEMM_MAX_STAGES=4 timeout 22 python bench.py --steps 10 --warmup 3 --latency-iters 5 \
  --cpu-queries 2 > gpurun_out/bench_ring4.json 2> gpurun_out/bench_ring4.err
echo "ring4 rc=$?"
ANR_GEMM_MAX_STAGES=3 timeout 13 python bench.py --steps 10 --warmup 3 --latency-iters 5 \
  --no-cpu-baseline > gpurun_out/bench_ring3.json 2> gpurun_out/bench_ring3.err
echo "ring3 rc=$?"
python - <<'PY'
import json
for f in ("bench_ring4", "bench_ring3"):
    try:
        d = json.load(open("gpurun_out/" + f + ".json"))
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "dense", round(d["roofline"]["avg_launch_ms"], 4),
              "bm25 in-step", round(d["roofline_other"].get("in_step_ms") or 0, 4),
              "bm25 alone", round(d["roofline_other"]["avg_launch_ms"], 4), "b1", round(d["batch1"]["device_ms"], 4),
              "parity", d.get("parity_checked_queries"), (d.get("cuda_graph") or {}).get("batch64"))
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
