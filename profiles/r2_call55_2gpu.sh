#!/bin/bash
# Round 2, GPU call 55 (2 GPUs): the full bench at N=2 as the driver launches it, after the
# programmatic-launch / in-place-query changes (parity of all 64 queries on the NCCL path).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T0=$SECONDS
timeout 600 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/c55_bench_${N}gpu.json 2> gpurun_out/c55_bench_${N}gpu.err
echo "bench N=$N rc=$? wall $((SECONDS - T0)) s"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c55_bench_${N}gpu.err | tail -5
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/c55_bench_{n}gpu.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "blocks", [round(x, 3) for x in d["blocks"]["ms_per_step_all"]], "parity", d["parity_checked_queries"], d.get("parity_error"))
print("multi_gpu", d["multi_gpu"])
print("e2e", round(d["e2e"]["value"]), "batch1", d["batch1"]["device_ms"])
for k, v in d["legs"].items():
    print("LEG", k, "parity", v.get("parity_checked_queries"), v.get("parity_error"), json.dumps(v)[:500])
PY
exit 0
