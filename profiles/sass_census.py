#!/usr/bin/env python
"""Static evidence that the kernels compile to the Blackwell paths DESIGN.md names: counts of the
tcgen05 / TMEM / TMA SASS mnemonics per kernel family in libanr_b200.so (cuobjdump -sass; runs
without a GPU).  Mnemonics as listed in B200_PROFILING.md: UTCHMMA = tcgen05.mma (f16/tf32 kinds),
.2CTA = cta_group::2, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc,
UTMALDG = cp.async.bulk.tensor (TMA tensor-map load, .MULTICAST across a cluster),
UBLKCP = cp.async.bulk (1-D TMA copy).

    python profiles/sass_census.py > profiles/r1_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "a-nice-rag_b200", "libanr_b200.so")
PAT = re.compile(r"\b((?:UTCHMMA|UTCQMMA|UTMALDG|UBLKCP|LDTM|UTCBAR|UTCATOMSWS)(?:\.\w+)*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)[1:]
    names = [p.split("\n", 1)[0].strip() for p in parts]
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    fam = collections.OrderedDict()
    for body, d in zip(parts, dem):
        d = d.replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", d).replace("void ", "").replace("anr::", "")
        family = re.sub(r"<.*", "", short)
        counts = collections.Counter(m.group(1) for m in PAT.finditer(body))
        n_inst = len(re.findall(r"/\*[0-9a-f]{4,6}\*/", body))
        f = fam.setdefault(family, {"variants": 0, "inst": [], "mn": collections.Counter()})
        f["variants"] += 1
        f["inst"].append(n_inst)
        for k, v in counts.items():
            f["mn"][k] = max(f["mn"][k], v)
    print("# SASS census of libanr_b200.so (sm_100a) -- `python profiles/sass_census.py`\n")
    print("Static instruction counts per kernel family (maximum over the template variants).\n")
    print("| kernel family | variants | SASS instructions | tcgen05 / TMEM / TMA mnemonics |")
    print("|---|---|---|---|")
    for family, f in fam.items():
        mn = ", ".join(f"{k} x{v}" for k, v in sorted(f["mn"].items())) or "-"
        lo, hi = min(f["inst"]), max(f["inst"])
        print(f"| `{family}` | {f['variants']} | {lo if lo == hi else f'{lo}-{hi}'} | {mn} |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
