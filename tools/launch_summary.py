#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total, mean, share."""
import csv
import sys
from collections import OrderedDict

src, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit() and r[-3] == "gpu__time_duration.sum"]
acc = OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    t = float(r[-1]) / 1e3 if r[-2] in ("ns", "nsecond") else float(r[-1])
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in acc.values())
print(cmd)
print("(cold-cache serialised times: compare SHARES with bench.py's kernels_ms_per_step)")
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-28s launches %3d  total %9.1f us  mean %9.1f us  share %.3f" % (name, n, t, t / n, t / tot))
