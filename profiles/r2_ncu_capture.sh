#!/bin/bash
# Round 2 ncu evidence (one GPU): launch list of a short bench run, then --set full captures of the
# two main kernels of a hybrid step as they run INSIDE the step (dense GEMM pass with the capped
# ring, BM25 pruned scan main launch).  Each ncu command follows the same command run plainly.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --blocks 1 --latency-iters 3 --no-cpu-baseline --legs headline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r2_bench_launches_ncu.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'dense_gemm_kernel|bm25_score_kernel|dense_tc_rescore' -s 40 -c 8 \
    -o gpurun_out/r2_full_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_bench_launches_ncu.csv
exit 0
