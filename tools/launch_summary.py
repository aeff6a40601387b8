#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"# {sys.argv[1]}: {sum(c for c, _ in agg.values())} launches, {tot/1e6:.3f} ms total (cold-cache, serialised: compare shares)")
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {c} | {t/1e3:.1f} | {100*t/tot:.1f}% | {t/c/1e3:.2f} |")
