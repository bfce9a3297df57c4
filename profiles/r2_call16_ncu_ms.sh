#!/bin/bash
# Round 2, GPU call 16: ncu --set full of the candidate-driven BM25 kernels (BM25-only batch of 64
# over 1M docs; the probe ran plainly first).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python profiles/bm25_probe.py 1000000 50000 64 > gpurun_out/c16_probe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ms_stage|ms_theta|ms_final' \
    -o gpurun_out/r2_ms_full python profiles/bm25_probe.py 1000000 50000 64 > gpurun_out/c16_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c16_ncu.log; ls -la gpurun_out/r2_ms_full.ncu-rep
