#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = [i for i, r in enumerate(rows[:10]) if "Line No" in r][0]
hdr = rows[hdr_i]
ls, samp, inst = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
exc = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
tot, lines = 0, []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[ls] == "" or r[2] != "-":
        continue
    s = int(float(r[samp] or 0))
    tot += s
    st = sorted(((h[6:], int(float(r[i] or 0))) for h, i in stalls), key=lambda kv: -kv[1])[:3]
    lines.append((int(r[ls]), r[1].strip()[:95], s, int(float(r[inst] or 0)), st, int(float(r[exc] or 0)) if exc else 0))
print("total samples", tot)
for l in sorted(lines, key=lambda x: -x[2])[:top_n]:
    print(f"{l[0]:4d} {100 * l[2] / max(tot, 1):5.1f}% inst={l[3]:>10d} exc_wf={l[5]:>10d} {l[1]:95s} {l[4]}")
