#!/bin/bash
# Round 2, GPU call 39: f4 (batched orchestrator) on the evaluator's shape, 1 and 2 dense models.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python profiles/f4_evaluator_bench.py --queries 2048 --models 1 > gpurun_out/c39_f4_m1.json 2> gpurun_out/c39_f4_m1.err
echo "m1 rc=$?"; tail -2 gpurun_out/c39_f4_m1.err; cat gpurun_out/c39_f4_m1.json | cut -c1-1200
timeout 500 python profiles/f4_evaluator_bench.py --queries 2048 --models 2 > gpurun_out/c39_f4_m2.json 2> gpurun_out/c39_f4_m2.err
echo "m2 rc=$?"; cat gpurun_out/c39_f4_m2.json | cut -c1-1200
