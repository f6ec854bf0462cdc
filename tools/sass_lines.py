#!/usr/bin/env python
"""Instruction count per source line of one kernel (nvdisasm --print-line-info on the cubin): where the code bytes go.

usage: sass_lines.py <nvdisasm output> <kernel name substring> [bucket boundaries as source line numbers ...]
"""
import collections
import re
import sys

path, kern = sys.argv[1], sys.argv[2]
bounds = [int(v) for v in sys.argv[3:]]
count = collections.Counter()
cur, inside = None, False
for ln in open(path):
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and cur:
        count[cur] += 1
total = sum(count.values())
print(f"{total} instructions, {total * 16 / 1024:.1f} kB")
if bounds:
    buckets = collections.Counter()
    for (f, l), c in count.items():
        if f != "backward.cu":
            buckets[f] += c
            continue
        k = sum(1 for b in bounds if l >= b)
        buckets[f"backward.cu[{bounds[k - 1] if k else 0}..{bounds[k] - 1 if k < len(bounds) else 'end'}]"] += c
    for k, c in sorted(buckets.items(), key=lambda kv: -kv[1]):
        print(f"{c:7d}  {c * 16 / 1024:6.1f} kB  {k}")
else:
    for (f, l), c in sorted(count.items(), key=lambda kv: -kv[1])[:60]:
        print(f"{c:7d}  {f}:{l}")
