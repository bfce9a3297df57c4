#!/bin/bash
# Round 2, GPU call 58: the content-word BM25 parity test (both paths) on the final library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "content_word" > gpurun_out/c58_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/c58_tests.log
