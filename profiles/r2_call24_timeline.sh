#!/bin/bash
# Round 2, GPU call 24: BM25 sub-timeline inside a hybrid step at 1M and at shard sizes; stage-1
# sample size knob.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local name=$1 n=$2; shift 2
  env "$@" timeout 200 python bench.py --chunks $n --steps 30 --warmup 5 --blocks 5 --latency-iters 5 --legs headline \
    --no-cpu-baseline > gpurun_out/c24_${name}.json 2> gpurun_out/c24_${name}.err
  echo "$name rc=$?"
}
run 1M 1000000 A=1
run 1M_s2048 1000000 ANR_MS_SAMPLE=2048
run 1M_s1024 1000000 ANR_MS_SAMPLE=1024
run 125k 125000 A=1
run 125k_s1024 125000 ANR_MS_SAMPLE=1024
run 125k_s512 125000 ANR_MS_SAMPLE=512
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c24_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        p = d.get("pipelined") or {}
        r, o = d["roofline"], d["roofline_other"]
        if r["kernel"].startswith("bm25"): r, o = o, r
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "bm25 alone", round(o.get("alone_ms") or 0, 4), "in-step", o.get("in_step_ms"),
              "graph", round(d["cuda_graph"]["batch64"]["replay_ms"], 4), "2inflight", round(p["two_in_flight"]["ms_per_step"], 4), "parity", d.get("parity_error"))
        print("   timeline", {k: v for k, v in d["timeline"].items() if k != "unit"})
    except Exception as e:
        print(f, "ERR", e)
PY
