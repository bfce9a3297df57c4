#!/bin/bash
# Round 2, GPU call 53: the tree as the driver will run it: smoke(), the GPU suite, the default
# bench, the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c53_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/c53_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c53_suite.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/c53_suite.log
T0=$SECONDS
timeout 600 python bench.py > gpurun_out/c53_bench.json 2> gpurun_out/c53_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -3 gpurun_out/c53_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/c53_ref.json 2> gpurun_out/c53_ref.err
echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c53_bench.json") if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/c53_ref.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "profiled", round(d["blocks"]["profiled"]["ms_per_step"], 4))
print("  e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"), "reruns", d.get("reruns", {}).get("dense_queries"), d.get("reruns", {}).get("bm25_queries"), "launches", d["gpu_launches"])
ro = d["roofline"]; o = d["roofline_other"]
print("  roofline", ro["kernel"], round(ro["avg_launch_ms"], 4), round(ro["frac"], 3), "alone", ro.get("alone_ms"), ro.get("alone_frac"), "| other in step", o.get("in_step_ms"), "alone", o.get("alone_ms"))
print("  graph", d["cuda_graph"]["batch64"], "2inflight", d["pipelined"]["two_in_flight"]["ms_per_step"], "batch1", d["batch1"]["device_ms"], "fp32 scan", d["batch1_fp32_scan"]["device_ms"], "filtered", d["filtered"]["ms_per_step"], "content", d["content_words"])
for k, v in d.get("legs", {}).items():
    print("  LEG", k, "parity", v.get("parity_checked_queries"), v.get("parity_error"), json.dumps(v)[:300])
print("  clocks", d["clocks"])
print("ref", r["value"], r["cpu_baseline"]["cores"], "same config", r["config"] == d["config"])
PY
