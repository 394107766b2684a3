#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean ms."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    tot[name][0] += 1
    tot[name][1] += v
total = sum(t for _, t in tot.values())
print("%-64s %7s %11s %10s %6s" % ("kernel", "count", "total ms", "mean ms", "share"))
for name, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-64s %7d %11.3f %10.4f %5.1f%%" % (name[:64], c, t, t / c, 100 * t / total))
