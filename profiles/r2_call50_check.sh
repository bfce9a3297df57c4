#!/bin/bash
# Round 2, GPU call 50: the new hybrid-with-reruns test, the dense / graph test files again (chain-final
# kernels no longer trigger early; profiled main kernel launched plainly), then the full default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_graph.py -x -q > gpurun_out/c50_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c50_tests.log
T0=$SECONDS
timeout 600 python bench.py > gpurun_out/c50_bench.json 2> gpurun_out/c50_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -3 gpurun_out/c50_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c50_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "profiled", round(d["blocks"]["profiled"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"), "reruns", d.get("reruns"))
r = d["roofline"]; o = d["roofline_other"]
print("  roofline", r["kernel"], r["avg_launch_ms"], r["frac"], "alone", r.get("alone_ms"), r.get("alone_frac"), "| other in step", o.get("in_step_ms"), "alone", o.get("alone_ms"))
print("  graph", d["cuda_graph"]["batch64"], "2inflight", d["pipelined"]["two_in_flight"]["ms_per_step"], "batch1", d["batch1"]["device_ms"], "fp32 scan", d["batch1_fp32_scan"]["device_ms"], "filtered", d["filtered"]["ms_per_step"])
for k, v in d.get("legs", {}).items():
    print("  LEG", k, "parity", v.get("parity_checked_queries"), v.get("parity_error"), json.dumps(v)[:420])
print("  clocks", d["clocks"])
PY
