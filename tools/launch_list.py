#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list (count, total, average,
share).  usage: launch_list.py launches.csv [title]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1000.0 if r[iu] in ("ns", "nsecond") else (v * 1000.0 if r[iu] in ("ms", "msecond") else v)
    acc[r[ik]][0] += 1
    acc[r[ik]][1] += us
tot = sum(v[1] for v in acc.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# %d launches captured, total %.3f ms (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)" % (sum(v[0] for v in acc.values()), tot / 1000))
for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f %%  %5d launches  %10.1f us total  %9.1f us avg   %s" % (100 * us / tot, n, us, us / n, k[:150]))
