#!/bin/bash
# Round 2, GPU call 20: candidate-driven BM25 with bucket tables + evaluation order: parity, then
# the headline, then the per-kernel durations of one BM25-only batch (ncu launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -k "bm25 or hybrid or config" > gpurun_out/c20_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/c20_tests.log
timeout 600 python bench.py --legs headline,big --blocks 7 > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/c20_bench.err
python profiles/bm25_probe.py 1000000 50000 64 > gpurun_out/c20_probe.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'ms_|topk_final|bm25_score' \
    --log-file gpurun_out/c20_bm25_launches.csv python profiles/bm25_probe.py 1000000 50000 64 > gpurun_out/c20_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json, csv
d = json.loads([l for l in open("gpurun_out/c20_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]],
      "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"))
print("  roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms", "alone_ms", "alone_frac")})
print("  other", {k: d["roofline_other"].get(k) for k in ("kernel", "avg_launch_ms", "in_step_ms", "alone_ms")})
print("  batch1", d["batch1"], "graph", d["cuda_graph"], "filtered", d["filtered"]["ms_per_step"])
print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
print("  pipelined", d.get("pipelined"))
for k, v in d.get("legs", {}).items():
    print("  LEG", k, json.dumps(v)[:700])
rows = [l for l in open("gpurun_out/c20_bm25_launches.csv") if l.startswith('"')]
for r in csv.DictReader(rows):
    print(r["Kernel Name"][:60], r["Grid Size"], r["Metric Value"], r["Metric Unit"])
PY
exit 0
