#!/bin/bash
# Round 2, GPU call 11 (one GPU): the records of the tree as committed -- the full default bench as
# the driver runs it (every leg at BASELINE size), the whole GPU suite, the reference arm, then the
# ncu launch list and --set full captures of the step's main kernels (each after the same command
# ran plainly and exited 0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 1200 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -3 gpurun_out/c11_bench.err
T0=$SECONDS
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/c11_suite.log 2>&1
echo "suite rc=$? wall $((SECONDS - T0)) s"; tail -25 gpurun_out/c11_suite.log
T0=$SECONDS
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/c11_ref.json 2> gpurun_out/c11_ref.err
echo "ref rc=$? wall $((SECONDS - T0)) s"
CMD="python bench.py --steps 3 --warmup 3 --blocks 1 --latency-iters 3 --no-cpu-baseline --legs headline"
$CMD > gpurun_out/c11_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r2_bench_launches_ncu.csv $CMD > gpurun_out/c11_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'dense_gemm_kernel|bm25_score|dense_tc_rescore' -s 40 -c 8 \
    -o gpurun_out/r2_full_step $CMD > gpurun_out/c11_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/c11_ncu_full.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_bench_launches_ncu.csv
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c11_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
      "parity", d["parity_checked_queries"], d.get("parity_error"))
for k, v in d.get("legs", {}).items():
    print("LEG", k, json.dumps(v)[:600])
PY
exit 0
