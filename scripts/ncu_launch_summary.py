#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares."""
import collections
import csv
import sys


def main(path, per_launch=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[row['Metric Unit']]
        name = row['Kernel Name']
        agg.setdefault(name, []).append(v)
        order.append((name, v, row.get('Grid Size', '')))
    tot = sum(sum(v) for v in agg.values())
    print(f'{"kernel":80s} {"n":>5s} {"total ms":>9s} {"share":>6s} {"avg us":>9s}')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k[:80]:80s} {len(v):5d} {sum(v) / 1e3:9.3f} {100 * sum(v) / tot:5.1f}% {sum(v) / len(v):9.1f}')
    print(f'total {tot / 1e3:.3f} ms over {sum(len(v) for v in agg.values())} launches')
    if per_launch:
        for name, v, g in order:
            print(f'{v:9.1f} us  grid {g:>14s}  {name[:70]}')


if __name__ == '__main__':
    main(sys.argv[1], len(sys.argv) > 2)
