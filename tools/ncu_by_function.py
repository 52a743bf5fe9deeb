#!/usr/bin/env python
"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by the ox_stages.cuh / ox_spec.cuh FUNCTION each
source line belongs to (samples, executed warp-instructions, static SASS instructions).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_by_function.py src.csv"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
cur_file, cur = None, None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r[0] == "Line No":
        iS, iI = r.index("# Samples"), r.index("Instructions Executed")
        continue
    if r[0] != "":
        cur = (cur_file, int(r[0]))
        continue
    if cur is None:
        continue
    f = lambda s: float(s) if s not in ("", "-") else 0.0
    a = agg[cur]
    a[0] += f(r[iS]); a[1] += f(r[iI]); a[2] += 1
funcs = {}
for fl in {k[0] for k in agg}:
    path = os.path.join(ROOT, "oxide_control_b200", "csrc", fl)
    lst = []
    if os.path.exists(path):
        for i, l in enumerate(open(path), 1):
            m = re.match(r"\s*(?:static |template <[^>]*> )*(?:OX_HDN?|__device__ __noinline__|__device__|__global__) .*?(\w+)\([^;]*$", l.rstrip())
            if m and "{" in l:
                lst.append((i, m.group(1)))
    funcs[fl] = lst


def fn(fl, ln):
    name = "?"
    for l, n in funcs.get(fl, ()):
        if l <= ln:
            name = n
        else:
            break
    return fl.split(".")[0] + ":" + name


tot = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (fl, ln), v in agg.items():
    k = fn(fl, ln)
    for i in range(3):
        tot[k][i] += v[i]
S, I, N = (sum(v[i] for v in tot.values()) for i in range(3))
print(f"samples {S:.0f}  executed warp-inst {I:.3e}  static inst {N:.0f}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{k:44s} samp {100 * v[0] / S:5.1f}%  executed {100 * v[1] / I:5.1f}%  static {100 * v[2] / N:5.1f}%")
