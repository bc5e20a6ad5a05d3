#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: summarize_launches.py <csv> [title]
Iterations in the capture = adam_kernel launches / 2 (one Adam update per network per training iteration)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
    rows.append((r['Kernel Name'], r['Grid Size'], ns))
short = lambda n: re.sub(r'^(void )?(b200gan::)?(\(anonymous namespace\)::)?', '', n).split('(')[0][:72]
agg = collections.OrderedDict()
for name, grid, ns in rows:
    a = agg.setdefault(short(name), [0, 0.0, set()])
    a[0] += 1; a[1] += ns; a[2].add(grid)
iters = max(1, sum(1 for n, _, _ in rows if 'adam_kernel' in n) // 2)
total = sum(ns for _, _, ns in rows)
print(f'{len(rows)} launches, {total / 1e6:.1f} ms of kernel time, {iters} training iterations in the capture = {total / 1e6 / iters:.2f} ms per iteration under ncu')
print()
print('| ms / iteration | share | launches / iter | ms each | kernel | grids |')
print('|---:|---:|---:|---:|---|---|')
for name, (cnt, ns, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| {ns / 1e6 / iters:.3f} | {100 * ns / total:.1f}% | {cnt / iters:.1f} | {ns / cnt / 1e6:.3f} | `{name}` | {", ".join(sorted(grids)[:3])} |')
