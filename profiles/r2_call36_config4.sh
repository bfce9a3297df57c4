#!/bin/bash
# Round 2, GPU call 36: where BASELINE configs[3] (BM25-only, 10M docs, V=500k, batch 256) spends its
# time in the candidate-driven path: per-query statistics + per-kernel durations; bench watchdog check.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ANR_MS_DEBUG=1 timeout 300 python profiles/bm25_probe.py 10000000 500000 256 2> gpurun_out/c36_stats.log > gpurun_out/c36_probe.log
grep "anr ms\] nq" gpurun_out/c36_stats.log | tail -2
python - <<'PY'
import re
rows = [l for l in open("gpurun_out/c36_stats.log") if "[anr ms]   q" in l]
rows = rows[-256:]
s2 = sorted(int(re.search(r"s2 (\d+)", l).group(1)) for l in rows)
sv = sorted(int(re.search(r"surv (\d+)", l).group(1)) for l in rows)
print("queries", len(rows), "s2 median", s2[len(s2)//2], "p90", s2[int(len(s2)*0.9)], "max", s2[-1], "sum", sum(s2))
print("survivors median", sv[len(sv)//2], "p90", sv[int(len(sv)*0.9)], "max", sv[-1])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'ms_|topk_final|bm25_score' \
    --log-file gpurun_out/c36_launches.csv python profiles/bm25_probe.py 10000000 500000 256 > gpurun_out/c36_ncu.log 2>&1
python - <<'PY'
import csv
rows = [l for l in open("gpurun_out/c36_launches.csv") if l.startswith('"')]
for r in list(csv.DictReader(rows))[-8:]:
    print(r["Kernel Name"][:50], r["Grid Size"], r["Metric Value"], r["Metric Unit"])
PY
timeout 200 python bench.py --steps 5 --warmup 3 --blocks 2 --latency-iters 3 --no-cpu-baseline --legs headline,big --leg-budget-s 2 \
    > gpurun_out/c36_watchdog.json 2> gpurun_out/c36_watchdog.err
echo "watchdog run rc=$?"; python -c "
import json; d = json.loads([l for l in open('gpurun_out/c36_watchdog.json') if l.startswith('{')][-1]); print('legs', d['legs'], 'value', round(d['value']))"
