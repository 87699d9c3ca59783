#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share."""
import collections
import csv
import re
import sys


def main(path, sequence=0):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, total, seq = collections.OrderedDict(), 0.0, []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}[row['Metric Unit']]
        name = re.sub(r'\(.*', '', row['Kernel Name'])[:90]
        seq.append((name, v, row.get('Grid Size', ''), row.get('Block Size', '')))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {total:.1f} us of kernel time '
          '(cold-cache, serialised by ncu: compare shares, not absolutes)')
    print('%10s %7s %6s %9s  %s' % ('total_us', 'share', 'n', 'avg_us', 'kernel'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%10.1f %6.1f%% %6d %9.1f  %s' % (t, 100 * t / total, n, t / n, k))


    if sequence:
        print(f'# the last {sequence} launches in order (us, grid, block, kernel)')
        for name, v, g, b in seq[-sequence:]:
            print('%9.1f  %-14s %-12s %s' % (v, g, b, name))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
