#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of device time)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in r:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(',', ''))
        except ValueError:
            continue
        if row[ui] == 'ns':
            v /= 1e3
        elif row[ui] == 'ms':
            v *= 1e3
        name = re.sub(r'\(.*', '', row[ki])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print('total_us %.1f launches %d' % (tot, sum(a[0] for a in agg.values())))
    print('%8s %12s %7s %10s  kernel' % ('launches', 'total_us', 'share', 'avg_us'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%8d %12.1f %6.1f%% %10.1f  %s' % (n, t, 100 * t / tot, t / n, k))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
