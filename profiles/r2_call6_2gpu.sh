#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): the full bench at N=2 as the driver launches it (parity of all 64
# queries on the real NCCL path, multi_gpu part timings, weak leg at 12.5M chunks per GPU), the
# reference arm under torchrun, and the captured sharded step (sharded_graph_probe.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T0=$SECONDS
timeout 1500 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/c6_bench_${N}gpu.json 2> gpurun_out/c6_bench_${N}gpu.err
echo "bench N=$N rc=$? wall $((SECONDS - T0)) s"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c6_bench_${N}gpu.err | tail -5
timeout 600 $TR --master-port 29502 profiles/sharded_graph_probe.py \
  > gpurun_out/c6_graph_probe_${N}gpu.json 2> gpurun_out/c6_graph_probe_${N}gpu.err
echo "probe rc=$?"; grep "^{" gpurun_out/c6_graph_probe_${N}gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c6_graph_probe_${N}gpu.err | tail -3
timeout 600 $TR --master-port 29503 bench.py --impl reference --gpus $N --steps 5 --warmup 1 \
  > gpurun_out/c6_ref_${N}gpu.json 2> gpurun_out/c6_ref_${N}gpu.err
echo "ref rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
line = [l for l in open(f"gpurun_out/c6_bench_{n}gpu.json") if l.startswith("{")][-1]
d = json.loads(line)
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "parity", d["parity_checked_queries"], d.get("parity_error"))
print("multi_gpu", d["multi_gpu"])
print("e2e", round(d["e2e"]["value"]), "batch1", d["batch1"]["device_ms"], "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"))
print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms")}, "other", {k: d["roofline_other"].get(k) for k in ("kernel", "in_step_ms")})
for k, v in d["legs"].items():
    print("LEG", k, json.dumps(v)[:1200])
r = json.loads([l for l in open(f"gpurun_out/c6_ref_{n}gpu.json") if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"]["cores"], "same config", r["config"] == d["config"])
PY
exit 0
