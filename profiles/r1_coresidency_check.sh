#!/bin/bash
# Co-residency experiment, second take: capped GEMM ring WITH the max-shared carveout preference.
cd "$(dirname "$0")/.."
for st in 4 3; do
  ANR_GEMM_MAX_STAGES=$st timeout 20 python bench.py --steps 10 --warmup 3 --latency-iters 5 \
    --no-cpu-baseline > gpurun_out/bench_ring$st.json 2> gpurun_out/bench_ring$st.err
  echo "ring$st rc=$?"
done
python - <<'PY'
import json
for f in ("bench_ring4", "bench_ring3"):
    try:
        d = json.load(open("gpurun_out/" + f + ".json"))
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "dense", round(d["roofline"]["avg_launch_ms"], 4),
              "bm25 in-step", round(d["roofline_other"].get("in_step_ms") or 0, 4),
              "bm25 alone", round(d["roofline_other"]["avg_launch_ms"], 4), "b1", round(d["batch1"]["device_ms"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
