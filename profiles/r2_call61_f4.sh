#!/bin/bash
# Round 2, GPU call 61: f4 after the vectorised code -> id-string step: its GPU parity test, then the
# evaluator-shaped timing (one dense model + BM25, 2048 queries, full ranking of 12 000 ids).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -k "retrieve_documents_batch" > gpurun_out/c61_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c61_tests.log
timeout 200 python profiles/f4_evaluator_bench.py --queries 2048 --models 1 --cpu-queries 3 > gpurun_out/c61_f4_m1.json 2> gpurun_out/c61_f4_m1.err
echo "m1 rc=$?"; tail -2 gpurun_out/c61_f4_m1.err; cut -c1-900 gpurun_out/c61_f4_m1.json
