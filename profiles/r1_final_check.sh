#!/bin/bash
# Last GPU call of round 1: the new known-answer test, a short default bench (checks the edited
# bench.py end to end) and the co-residency experiment.
cd "$(dirname "$0")/.."
timeout 25 python -m pytest tests/test_gpu_parity.py -q -k "known_answer" > gpurun_out/kat_test.log 2>&1
echo "rc=$?" >> gpurun_out/kat_test.log
timeout 40 python bench.py --steps 10 --warmup 3 --latency-iters 30 --cpu-queries 2 \
  > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
echo "bench rc=$?"
ANR_GEMM_MAX_STAGES=4 timeout 30 python bench.py --steps 10 --warmup 3 --latency-iters 10 \
  --no-cpu-baseline > gpurun_out/bench_stages4.json 2> gpurun_out/bench_stages4.err
echo "bench4 rc=$?"
tail -3 gpurun_out/kat_test.log
python - <<'PY'
import json
for f in ("bench_r1f", "bench_stages4"):
    try:
        d = json.load(open("gpurun_out/" + f + ".json"))
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "dense", round(d["roofline"]["avg_launch_ms"], 4),
              "bm25 in-step", d["roofline_other"].get("in_step_ms"), "b1", round(d["batch1"]["device_ms"], 4),
              d.get("cuda_graph"), "parity", d.get("parity_checked_queries"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_r1f.err gpurun_out/bench_stages4.err
