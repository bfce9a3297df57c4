#!/bin/bash
# Round 2, GPU call 52: profiles/pdl_probe.py -- the step under every "pdl" mask, interleaved.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/pdl_probe.py > gpurun_out/c52_pdl_probe.json 2> gpurun_out/c52_pdl_probe.err
echo "probe rc=$?"; tail -3 gpurun_out/c52_pdl_probe.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c52_pdl_probe.json") if l.startswith("{")][-1])
for b, r in d["batches"].items():
    for m, v in r.items():
        print("batch", b, "pdl", m, round(v["median"], 4), [round(x, 4) for x in v["all"]], v["identical_to_serialised"])
PY
