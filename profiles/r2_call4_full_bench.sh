#!/bin/bash
# Round 2, GPU call 4: whole GPU suite, then the full default bench (every leg at BASELINE size)
# and the reference arm, as the driver runs them.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 1200 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -5 gpurun_out/c4_bench.err
T0=$SECONDS
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/c4_ref.json 2> gpurun_out/c4_ref.err
echo "ref rc=$? wall $((SECONDS - T0)) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/c4_bench.json"))
r = json.load(open("gpurun_out/c4_ref.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 4) for x in d["blocks"]["ms_per_step_all"]])
print("e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"))
print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms", "alone_ms", "alone_frac", "traffic")})
print("other", {k: d["roofline_other"].get(k) for k in ("kernel", "frac", "avg_launch_ms", "in_step_ms", "alone_ms", "traffic")})
print("cpu", d.get("cpu_baseline"))
print("batch1", d["batch1"], d["batch1_fp32_scan"])
print("graph", d["cuda_graph"], "filtered", d["filtered"])
print("clocks", d["clocks"])
for k, v in d["legs"].items():
    print("LEG", k, json.dumps(v)[:1500])
print("ref", r["value"], r["cpu_baseline"], r["config"] == d["config"])
PY
exit 0
