#!/bin/bash
# Round 2, GPU call 3: BM25 scan by runs (cursors kept across tiles) -- parity suite, then timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c3_suite.log 2>&1
echo "suite rc=$?"; tail -15 gpurun_out/c3_suite.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --blocks 3 --latency-iters 5 \
    --no-cpu-baseline --legs headline \
    > gpurun_out/c3_${name}.json 2> gpurun_out/c3_${name}.err
  echo "$name rc=$?"
}
run default A=1
run oldkernel ANR_BM25_RUNS=0
run run2 ANR_BM25_RUN_TILES=2
run run8 ANR_BM25_RUN_TILES=8
run tile4096_run6 ANR_BM25_TILE=4096 ANR_BM25_RUN_TILES=6
run tile3072_run8 ANR_BM25_TILE=3072 ANR_BM25_RUN_TILES=8
run div4 ANR_BM25_HEAD_DIV=4
run div4_ring3 ANR_BM25_HEAD_DIV=4 ANR_GEMM_BESIDE_STAGES=3
run nobeside ANR_GEMM_BESIDE_STAGES=0
# the whole new bench.py once at reduced sizes (every leg, parity of every query): a functional check
timeout 600 python bench.py --steps 5 --warmup 3 --blocks 2 --latency-iters 5 --chunks 200000 \
  --big-chunks 400000 --big-vocab 50000 --weak-chunks-per-gpu 300000 --cpu-queries 8 --leg-steps 3 \
  > gpurun_out/c3_small_full.json 2> gpurun_out/c3_small_full.err
echo "small full bench rc=$?"; tail -5 gpurun_out/c3_small_full.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c3_small_full.json"))
    print("small full:", round(d["value"]), "parity", d["parity_checked_queries"], d.get("parity_error"),
          {k: (v.get("parity_checked_queries"), v.get("parity_error"), v.get("error")) if isinstance(v, dict) else v
           for k, v in d["legs"].items()})
except Exception as e:
    print("small full ERR", e)
PY
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c3_*.json")):
    if "small_full" in f: continue
    try:
        d = json.load(open(f))
        ro, rd = d["roofline_other"], d["roofline"]
        if rd["kernel"].startswith("bm25"): ro, rd = rd, ro
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), "dense", round(rd["avg_launch_ms"], 4),
              "bm25 in-step", round(ro.get("in_step_ms") or 0, 4),
              "bm25 alone", round(ro["avg_launch_ms"], 4), "b1", round(d["batch1"]["device_ms"], 4),
              "e2e", round(d["e2e"]["value"]), "graph", round((d.get("cuda_graph") or {}).get("batch64",{}).get("replay_ms",0),4))
    except Exception as e:
        print(f, "ERR", e)
PY
exit 0
