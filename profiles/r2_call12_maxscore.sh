#!/bin/bash
# Round 2, GPU call 12: candidate-driven BM25 top-k (anr_bm25_ms.cu) -- parity first (BM25 tests,
# memcheck on the new test, the whole suite), then the headline and the big legs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bm25" > gpurun_out/c12_bm25_tests.log 2>&1
echo "bm25 tests rc=$?"; tail -30 gpurun_out/c12_bm25_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "candidate_path or pruned_with_doc_mask" > gpurun_out/c12_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/c12_memcheck.log | tail -8
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c12_suite.log 2>&1
echo "suite rc=$?"; tail -12 gpurun_out/c12_suite.log
timeout 900 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/c12_bench.err
ANR_BM25_MAXSCORE=0 timeout 300 python bench.py --legs headline --no-cpu-baseline --blocks 5 > gpurun_out/c12_bench_tiled.json 2> gpurun_out/c12_bench_tiled.err
python - <<'PY'
import json
for f in ("gpurun_out/c12_bench.json", "gpurun_out/c12_bench_tiled.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]],
          "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"))
    print("  roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms", "alone_ms", "alone_frac")})
    print("  other", {k: d["roofline_other"].get(k) for k in ("kernel", "avg_launch_ms", "in_step_ms", "alone_ms")})
    print("  batch1", d["batch1"], "graph", d["cuda_graph"], "filtered", d["filtered"]["ms_per_step"])
    print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
    print("  pipelined", d.get("pipelined"))
    for k, v in d.get("legs", {}).items():
        print("  LEG", k, json.dumps(v)[:900])
PY
exit 0
