#!/bin/bash
# Round 2, GPU call 51: ms_theta without the early trigger (stage 2 no longer parks on the SMs ahead
# of the dense main kernel); the hybrid-with-reruns test; headline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -x -q -k "reruns_on_both or overflow" > gpurun_out/c51_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c51_tests.log
timeout 200 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --cpu-queries 8 \
    > gpurun_out/c51_bench.json 2> gpurun_out/c51_bench.err
echo "bench rc=$?"
python - <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/c51_bench.json") if l.startswith("{")][-1])
p = d.get("pipelined") or {}
r = d["roofline"]
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "profiled", round(d["blocks"]["profiled"]["ms_per_step"], 4), "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_checked_queries"), d.get("parity_error"))
print("  dense kernel in step", r.get("avg_launch_ms"), "alone", r.get("alone_ms"), "bm25 in step", d["roofline_other"].get("in_step_ms"), "e2e", round(d["e2e"]["value"]), "sync e2e", round(d["e2e"]["synchronous_call"]["value"]))
print("  batch1", d["batch1"]["device_ms"], "clocks", d["clocks"]["sm_mhz"])
print("  timeline", {k: v for k, v in d["timeline"].items() if k not in ("unit",)})
PY
