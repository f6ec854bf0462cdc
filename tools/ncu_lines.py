#!/usr/bin/env python
"""Samples and executed instructions per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.  usage: ncu_lines.py file.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, ""])  # (file, line) -> samples, executed
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    smp, ex = r[hdr["# Samples"]], r[hdr["Instructions Executed"]]
    if not ex.isdigit():
        continue
    e = agg[(cur_file, int(r[0]))]
    e[0] += int(smp) if smp.isdigit() else 0
    e[1] += int(ex)
    e[2] = r[1].strip()[:90]
tot_s = sum(v[0] for v in agg.values())
tot_e = sum(v[1] for v in agg.values())
print(f"total samples {tot_s}  executed {tot_e}")
byfile = defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]
    byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:20s} samples {100 * v[0] / max(tot_s, 1):5.1f}%  executed {100 * v[1] / max(tot_e, 1):5.1f}%")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / max(tot_s, 1):5.1f}% smp {100 * v[1] / max(tot_e, 1):5.1f}% ex  {f}:{l:<4d} {v[2]}")
