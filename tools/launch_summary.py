#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [tiles]"""
import collections
import csv
import sys

tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 'usecond': v, 'nsecond': v / 1e3, 'msecond': v * 1e3}.get(row['Metric Unit'], v)
    name = row['Kernel Name'][:78]
    agg.setdefault(name, [0.0, 0])
    agg[name][0] += v
    agg[name][1] += 1
    tot += v
print(f"{'us/tile':>10s} {'launches':>8s} {'share':>6s}  kernel")
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v / tiles:10.1f} {c:8d} {100 * v / tot:5.1f}%  {k}")
print(f"total {tot / tiles:.1f} us per tile (cold-cache, serialised: compare shares, not absolutes)")
