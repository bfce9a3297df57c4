#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples from `ncu --page source --csv --print-source
cuda,sass` output.  usage: ncu_source_lines.py <csv> <result-index> [top-n]"""
import csv, collections, io, sys
path, which = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lines = open(path).read().splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Function Name"')]
files = [lines[s - 1].split('","')[1].rstrip('"') for s in starts]
# results = groups of sections, a new group starts whenever the first file repeats
groups, cur = [], []
for s, f in zip(starts, files):
    if cur and f == files[0]:
        groups.append(cur); cur = []
    cur.append((s, f))
groups.append(cur)
g = groups[which]
ends = {s: (starts[starts.index(s) + 1] - 1 if starts.index(s) + 1 < len(starts) else len(lines)) for s in starts}
per_line = collections.defaultdict(lambda: [0, 0, ""])
tot_i = tot_s = 0
for s, f in g:
    rd = csv.reader(io.StringIO("\n".join(lines[s + 1:ends[s]])))
    hdr = next(rd)
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    line_no, src = None, ""
    for r in rd:
        if len(r) < len(hdr): continue
        if r[0]:
            line_no, src = r[0], r[1]
        try:
            ie, sm = int(r[iI] or 0), int(r[iS] or 0)
        except ValueError:
            continue
        key = (f.split("/")[-1], int(line_no))
        per_line[key][0] += ie; per_line[key][1] += sm; per_line[key][2] = src
        tot_i += ie; tot_s += sm
print(f"total warp instructions {tot_i}, samples {tot_s}")
for key, (ie, sm, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*ie/max(tot_i,1):5.1f}% inst {100*sm/max(tot_s,1):5.1f}% samp  {key[0]}:{key[1]:<5d} {src.strip()[:110]}")
