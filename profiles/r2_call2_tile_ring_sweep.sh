#!/bin/bash
# Round 2, GPU call 2: BM25 tile size x dense ring depth under co-residency (how many BM25 CTAs
# fit beside the dense CTA), head-term threshold variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --latency-iters 5 --no-cpu-baseline \
    > gpurun_out/c2_${name}.json 2> gpurun_out/c2_${name}.err
  echo "$name rc=$?"
}
for ring in 3 4; do for tile in 3072 4096 5120; do
  run ring${ring}_tile${tile} ANR_GEMM_BESIDE_STAGES=$ring ANR_BM25_TILE=$tile
done; done
run ring4_div4 ANR_GEMM_BESIDE_STAGES=4 ANR_BM25_HEAD_DIV=4
run ring4_div16 ANR_GEMM_BESIDE_STAGES=4 ANR_BM25_HEAD_DIV=16
run ring3_tile4096_div16 ANR_GEMM_BESIDE_STAGES=3 ANR_BM25_TILE=4096 ANR_BM25_HEAD_DIV=16
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c2_*.json")):
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
