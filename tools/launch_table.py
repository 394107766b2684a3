#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total device time per kernel."""
import csv, sys, collections
tot = collections.OrderedDict()
cnt = collections.Counter()
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 10 and r[0].isdigit():
        name = r[4].split("(")[0].replace("void ", "").replace("b2k::", "")
        tot[name] = tot.get(name, 0) + int(r[-1])
        cnt[name] += 1
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if k.startswith("at::"): continue
    print("%-42s launches %4d  total %10.1f us  avg %10.1f us" % (k[:42], cnt[k], v / 1e3, v / 1e3 / cnt[k]))
