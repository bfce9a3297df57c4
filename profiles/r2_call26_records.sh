#!/bin/bash
# Round 2, GPU call 26: records of the tree with the candidate-driven BM25 path: whole GPU suite,
# full default bench, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c26_suite.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/c26_suite.log
T0=$SECONDS
timeout 1200 python bench.py > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -3 gpurun_out/c26_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/c26_ref.json 2> gpurun_out/c26_ref.err
echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c26_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]],
      "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"))
print("  roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms", "alone_ms", "alone_frac", "frac_traffic")})
print("  other", {k: d["roofline_other"].get(k) for k in ("avg_launch_ms", "in_step_ms", "alone_ms", "frac", "frac_traffic")})
print("  batch1", d["batch1"], d["batch1_fp32_scan"], "graph", d["cuda_graph"], "filtered", d["filtered"]["ms_per_step"])
print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
print("  pipelined", d.get("pipelined"))
print("  cpu", d.get("cpu_baseline"))
for k, v in d.get("legs", {}).items():
    print("  LEG", k, json.dumps(v)[:1000])
PY
