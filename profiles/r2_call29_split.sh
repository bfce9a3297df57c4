#!/bin/bash
# Round 2, GPU call 29: which L1 split for the early BM25 kernels (plan / stage 1 / theta) of a
# hybrid step: the dense kernels' (default) or the default split (ANR_MS_EARLY_SPLIT=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local name=$1 n=$2; shift 2
  env "$@" timeout 200 python bench.py --chunks $n --steps 30 --warmup 5 --blocks 7 --latency-iters 5 --legs headline \
    --no-cpu-baseline > gpurun_out/c29_${name}.json 2> gpurun_out/c29_${name}.err
  echo "$name rc=$?"
}
run 1M_all_shared 1000000 A=1
run 1M_early_default 1000000 ANR_MS_EARLY_SPLIT=0
run 125k_all_shared 125000 A=1
run 125k_early_default 125000 ANR_MS_EARLY_SPLIT=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c29_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        p = d.get("pipelined") or {}
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
        print("   timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
    except Exception as e:
        print(f, "ERR", e)
PY
