#!/bin/bash
# Round 2, GPU call 28: dense tests after the rescoring-kernel change, the headline, then the ncu
# evidence of the final kernels: launch list of a short bench run and --set full captures of the
# kernels of one hybrid step (dense GEMM pass, candidate-driven BM25 chain, rescoring).  Each ncu
# command follows the same command run plainly.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q > gpurun_out/c28_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c28_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 --blocks 7 --latency-iters 20 --legs headline --no-cpu-baseline \
    > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err
echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --blocks 1 --latency-iters 3 --no-cpu-baseline --legs headline"
$CMD > gpurun_out/c28_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2b_bench_launches_ncu.csv $CMD > gpurun_out/c28_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'dense_gemm_kernel|ms_|dense_tc_rescore|dense_gemm_thr' -s 60 -c 9 \
    -o gpurun_out/r2b_full_step $CMD > gpurun_out/c28_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/c28_ncu_full.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c28_bench.json") if l.startswith("{")][-1])
p = d.get("pipelined") or {}
print("value", round(d["value"]), round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
print("  timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
print("  batch1", d["batch1"])
PY
