#!/bin/bash
# Round 2, GPU call 38: stage 2 keeps the looked-up weights (no second lookup pass for survivors).
cd "$(dirname "$0")/.."
bash profiles/r2_call27_stage2.sh d
timeout 200 python bench.py --chunks 125000 --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --no-cpu-baseline > gpurun_out/c38_125k.json 2> gpurun_out/c38_125k.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c38_125k.json") if l.startswith("{")][-1])
print("125k:", round(d["value"]), round(d["ms_per_step"], 4), "parity", d.get("parity_error"))
print("   timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
PY
