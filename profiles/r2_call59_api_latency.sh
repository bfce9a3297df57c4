#!/bin/bash
# Round 2, GPU call 59: latency of the drop-in Python API (SearchEngine calls on a 1M-row DataFrame,
# pandas result assembly included) on the final library: profiles/api_latency.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python profiles/api_latency.py 1000000 > gpurun_out/c59_api_latency.json 2> gpurun_out/c59_api_latency.err
echo "rc=$?"; tail -2 gpurun_out/c59_api_latency.err; grep "^{" gpurun_out/c59_api_latency.json
