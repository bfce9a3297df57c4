#!/bin/bash
# Round 2, GPU call 41: final check of the tree as the driver will run it: smoke(), the GPU suite,
# the default bench, the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c41_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/c41_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c41_suite.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/c41_suite.log
T0=$SECONDS
timeout 1200 python bench.py > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -3 gpurun_out/c41_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/c41_ref.json 2> gpurun_out/c41_ref.err
echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c41_bench.json") if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/c41_ref.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"), "reruns", d.get("reruns"))
print("  graph", d["cuda_graph"]["batch64"], "2inflight", d["pipelined"]["two_in_flight"]["ms_per_step"])
for k, v in d.get("legs", {}).items():
    print("  LEG", k, "parity", v.get("parity_checked_queries"), v.get("parity_error"), json.dumps(v)[:200])
print("ref", r["value"], r["cpu_baseline"]["cores"], "same config", r["config"] == d["config"])
PY
