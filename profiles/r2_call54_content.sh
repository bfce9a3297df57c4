#!/bin/bash
# Round 2, GPU call 54: profiles/content_words_probe.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/content_words_probe.py > gpurun_out/c54_content.json 2> gpurun_out/c54_content.err
echo "probe rc=$?"; tail -3 gpurun_out/c54_content.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c54_content.json") if l.startswith("{")][-1])
for k, v in d["sets"].items():
    print(k, json.dumps(v))
PY
