#!/bin/bash
# Round 2, GPU call 46: bf16 shadow copy in tiled order (16 KB blocks): whole GPU suite, then the
# headline with the tiled (default) and the row-major (ANR_SHADOW_TILED=0) copy on the same box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c46_suite.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/c46_suite.log
for T in 1 0; do
ANR_SHADOW_TILED=$T timeout 300 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --no-cpu-baseline \
    > gpurun_out/c46_bench_t$T.json 2> gpurun_out/c46_bench_t$T.err
echo "bench tiled=$T rc=$?"
python - $T <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/c46_bench_t%s.json" % sys.argv[1]) if l.startswith("{")][-1])
p = d.get("pipelined") or {}
r = d["roofline"]
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
print("  dense kernel in step", r.get("avg_launch_ms"), "alone", r.get("alone_ms"), "frac", r.get("frac"), r.get("alone_frac"))
print("  batch1", d["batch1"]["device_ms"], "clocks", d["clocks"])
PY
done
