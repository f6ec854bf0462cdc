#!/usr/bin/env python
"""Opcode mix and hot instructions of one kernel from `ncu -i X.ncu-rep --page source --csv [--launch-skip k --launch-count 1]`.
usage: ncu_source_mix.py file.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][:2])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
src, ex, smp = ix["Source"], ix["Instructions Executed"], ix["# Samples"]
mix, samp = defaultdict(int), defaultdict(int)
total = tsamp = 0
body = []
for r in rows[2:]:
    if len(r) <= ex or not r[ex].isdigit():
        continue
    n = int(r[ex])
    s = int(r[smp] or 0)
    text = r[src].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", text)
    op = m.group(2) if m else text[:12]
    base = op.split(".")[0]
    mix[base] += n
    samp[base] += s
    total += n
    tsamp += s
    body.append((n, s, r[ix["Address"]], text))
print(f"static instructions {len(body)}  executed {total}  samples {tsamp}")
for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k:14s} {v:14d} {100 * v / total:6.2f}%   samples {100 * samp[k] / max(tsamp, 1):6.2f}%")
print("---- hottest by samples")
for n, s, addr, text in sorted(body, key=lambda b: -b[1])[:top]:
    print(f"{s:8d} {n:12d}  {text[:110]}")
