#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and
share per kernel.  usage: launch_list.py <launches.csv> <title>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v if u.startswith("m") else v * 1e3
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print("# " + sys.argv[2])
print("# kernel, launches, total ms, share of all GPU time in the run")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-95s %6d %12.3f ms %6.2f%%" % (k[:95], a[0], a[1], 100 * a[1] / tot))
