#!/bin/bash
# Round 2, GPU calls 48-49: programmatic dependent launch along both kernel chains of a step (ANR_PDL,
# default on), weights of the step's fusion as kernel arguments, device-resident queries read in
# place: whole GPU suite, then the headline with and without PDL on the same box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/c49_suite.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/c49_suite.log
for P in 1 0; do
ANR_PDL=$P timeout 200 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --cpu-queries 8 \
    > gpurun_out/c49_bench_p$P.json 2> gpurun_out/c49_bench_p$P.err
echo "bench pdl=$P rc=$?"
python - $P <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/c49_bench_p%s.json" % sys.argv[1]) if l.startswith("{")][-1])
p = d.get("pipelined") or {}
r = d["roofline"]
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), d["cuda_graph"]["batch64"]["identical_to_eager"], "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_checked_queries"), d.get("parity_error"))
print("  dense kernel in step", r.get("avg_launch_ms"), "alone", r.get("alone_ms"), "e2e", round(d["e2e"]["value"]), "sync e2e", round(d["e2e"]["synchronous_call"]["value"]))
print("  batch1", d["batch1"]["device_ms"], d["cuda_graph"]["batch1"], "clocks", d["clocks"]["sm_mhz"])
print("  timeline", {k: v for k, v in d["timeline"].items() if k not in ("unit",)})
PY
done
