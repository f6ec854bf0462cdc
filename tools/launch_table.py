#!/usr/bin/env python
"""Per-iteration table of one solve from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_table.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dpilqr::", "")
    rows.append((name, int(row["Grid Size"].strip("()").split(",")[0]), int(row["Block Size"].strip("()").split(",")[0]),
                 float(row["Metric Value"].replace(",", "")) / 1e6))
tot = defaultdict(lambda: [0, 0.0])
for name, grid, blk, ms in rows:
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} {v[0]:5d} launches {v[1]:9.3f} ms {100 * v[1] / total:5.1f}%")
print(f"{'total':48s} {len(rows):5d} launches {total:9.3f} ms")
print("iter n_act | linquad backward | line-search launches: ms(grid x block)")
it, cur = 0, None
for name, grid, blk, ms in rows:
    if name.startswith("linquad"):
        cur = {"lq": ms, "ls": []}
    elif cur is None:
        continue
    elif name.startswith("backward"):
        cur["bw"], cur["n"] = ms, grid
    elif name.startswith("rollout_kernel"):
        cur["ls"].append(f"{ms:.2f}({grid}x{blk})")
    elif name.startswith("select_kernel"):
        it += 1
        print(f"{it:3d} {cur.get('n', 0):5d} | {cur['lq']:6.2f} {cur.get('bw', 0):7.2f} | " + "  ".join(cur["ls"]))
