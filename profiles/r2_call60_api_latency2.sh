#!/bin/bash
# Round 2, GPU call 60: drop-in layer after the registry probe / result-frame changes: the drop-in GPU
# tests, then profiles/api_latency.py again.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "dropin or similarity_search or retrieve_documents or frame" > gpurun_out/c60_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c60_tests.log
timeout 240 python profiles/api_latency.py 1000000 > gpurun_out/c60_api_latency.json 2> gpurun_out/c60_api_latency.err
echo "rc=$?"; tail -2 gpurun_out/c60_api_latency.err; grep "^{" gpurun_out/c60_api_latency.json
