#!/bin/bash
# Round 2, GPU call 56: ncu evidence of the final kernels: launch list of a short bench run
# (gpu__time_duration.sum, --clock-control none) and a --set full capture of one hybrid step's
# kernels (dense GEMM pass on the tiled bf16 shadow, thresholds, rescoring + flag list, BM25 chain,
# pair fusion).  Each ncu command follows the same command run plainly.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --blocks 1 --latency-iters 3 --no-cpu-baseline --legs headline"
$CMD > gpurun_out/c56_plain.log 2>&1
echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2c_bench_launches_ncu.csv $CMD > gpurun_out/c56_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'dense_gemm_kernel|ms_|dense_tc_rescore|dense_gemm_thr|wrrf_fuse_pair|f32_to_bf16' -s 60 -c 11 \
    -o gpurun_out/r2c_full_step $CMD > gpurun_out/c56_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/c56_ncu_full.log
ls -la gpurun_out/r2c_*
