#!/usr/bin/env python
"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_by_line.py src.csv [topN]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None
cur_file, cur = None, None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        iL = hdr.index("stall_long_sb")
        continue
    if r[0] != "":  # a CUDA source line row
        cur = (cur_file, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0.0, 0.0, 0.0, 0.0])
        continue
    if cur is None:
        continue
    a = agg[cur]
    f = lambda s: float(s) if s not in ("", "-") else 0.0
    a[0] += f(r[iS]); a[1] += f(r[iI]); a[2] += f(r[iT]); a[3] += f(r[iL])
tot = [sum(v[k] for v in agg.values()) for k in range(4)]
print(f"total samples {tot[0]:.0f}  warp-inst {tot[1]:.3e}  thread-inst {tot[2]:.3e}  long_sb samples {tot[3]:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:<5} samp {100 * v[0] / max(tot[0], 1):5.1f}%  inst {100 * v[1] / max(tot[1], 1):5.1f}%  lanes {v[2] / max(v[1], 1):4.1f}  longsb {100 * v[3] / max(v[0], 1):4.0f}%  {key[2][:110]}")
