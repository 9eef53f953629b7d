"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) per kernel: launches, average
duration and share of the summed kernel time.  usage: python tools/launch_list_summary.py X.csv "header line(s)" > out.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
t = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][-60:]
    t[name][0] += 1
    t[name][1] += float(r[-1]) / 1e3
total = sum(v[1] for v in t.values())
if len(sys.argv) > 2:
    print(sys.argv[2].replace("\\n", "\n"))
print(f"{'kernel':60s} {'launches':>10s} {'avg us':>12s} {'share':>8s}")
for k, (n, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {n:10d} {us / n:12.2f} {100 * us / total:7.2f}%")
