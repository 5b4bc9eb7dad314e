#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel for one step.
usage: launch_summary.py launches.csv [step_index] [marker]   (marker: kernel that starts a step, default prep_input)"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    marker = sys.argv[3] if len(sys.argv) > 3 else 'prep_input'
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [r['Kernel Name'] for r in rows]
    starts = [i for i, n in enumerate(names) if marker in n and (i == 0 or marker not in names[i - 1])]
    s = starts[step]
    e = starts[step + 1] if step + 1 < len(starts) else len(rows)
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[s:e]:
        n = r['Kernel Name'].split('(')[0].replace('<unnamed>::', '')
        t = float(r['Metric Value'].replace(',', ''))
        t *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r['Metric Unit'], 1.0)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print('step %d: launches %d..%d (%d kernels), summed device time %.1f us' % (step, s, e, e - s, tot))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-50s %4d launches %12.1f us %5.1f%%' % (n[:50], c, t, 100 * t / tot))


if __name__ == '__main__':
    main()
