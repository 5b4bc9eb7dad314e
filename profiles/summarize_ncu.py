#!/usr/bin/env python
"""Turns Nsight Compute output into the small tracked summaries under profiles/.

    # a `--metrics gpu__time_duration.sum --csv --log-file launches.csv` launch list -> per-kernel share table of ONE step
    python profiles/summarize_ncu.py launches gpurun_out/launches.csv --first 715 --count 213 > profiles/r02/launches_x_summary.txt

    # a `--set full` capture (.ncu-rep) -> raw csv + per-kernel / per-launch DRAM traffic, pipe utilisation (bench.py reads the json)
    python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep --out profiles/r02/ncu_full_train64 --note "..."

Per-launch times of an ncu run are cold-cache and serialised: compare SHARES with the CUDA-event numbers of bench.py, not absolutes.
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys


def short_name(k):
    k = re.sub(r'^void\s+', '', k)
    k = re.sub(r'<unnamed>::', '', k)
    k = re.sub(r'\(anonymous namespace\)::', '', k)
    m = re.match(r'([A-Za-z0-9_:]+(?:<[^(]*>)?)', k)
    return m.group(1) if m else k


def read_launch_csv(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(io.StringIO(''.join(lines)))
    header = next(rd)
    col = {n: i for i, n in enumerate(header)}
    for r in rd:
        if len(r) != len(header) or not r[col['ID']].isdigit():
            continue
        rows.append(r)
    return header, col, rows


def cmd_launches(a):
    header, col, rows = read_launch_csv(a.path)
    # long format: one row per (launch, metric)
    recs = {}
    for r in rows:
        if r[col['Metric Name']] != 'gpu__time_duration.sum':
            continue
        unit, val = r[col['Metric Unit']], float(r[col['Metric Value']].replace(',', ''))
        us = val / 1e3 if unit in ('ns', 'nsecond') else (val * 1e3 if unit in ('ms', 'msecond') else val)
        recs[int(r[col['ID']])] = (short_name(r[col['Kernel Name']]), us)
    ids = sorted(recs)
    first = a.first if a.first is not None else ids[0]
    sel = [i for i in ids if first <= i < first + (a.count or 10 ** 9)]
    tot = sum(recs[i][1] for i in sel)
    agg = {}
    for i in sel:
        n, us = recs[i]
        g = agg.setdefault(n, [0, 0.0])
        g[0] += 1
        g[1] += us
    print('launches %d..%d (%d kernels), summed device time %.1f us' % (sel[0], sel[-1], len(sel), tot))
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-52s %3d launches %12.1f us %5.1f%%' % (n[:52], c, us, 100 * us / tot))


WANT = {
    'time_us': 'gpu__time_duration.sum', 'dram_read': 'dram__bytes_read.sum', 'dram_write': 'dram__bytes_write.sum',
    'tensor_pipe_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_pipe_pct2': 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'fma_pipe_pct': 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm_busy_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'regs': 'launch__registers_per_thread',
}
UNIT_BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
UNIT_US = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}


def cmd_full(a):
    raw = subprocess.run(['ncu', '-i', a.path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    if a.out:
        open(a.out + '_raw.csv', 'w').write(raw)
    rd = csv.reader(io.StringIO(raw))
    header = next(rd)
    units = next(rd)
    idx = {}
    for key, metric in WANT.items():
        for i, h in enumerate(header):
            if h == metric or h.endswith('.' + metric):
                idx.setdefault(key, i)
    launches, per = [], {}
    for r in rd:
        if not r or not r[0].isdigit():
            continue
        rec = {'id': int(r[0]), 'kernel': short_name(r[header.index('Kernel Name')])}
        for key, i in idx.items():
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            u = units[i]
            if key.startswith('dram_r') or key.startswith('dram_w'):
                v *= UNIT_BYTES.get(u, 1.0)
            elif key == 'time_us':
                v *= UNIT_US.get(u, 1.0)
            rec[key] = v
        rec['dram_bytes'] = rec.pop('dram_read', 0.0) + rec.pop('dram_write', 0.0)
        if 'tensor_pipe_pct' not in rec and 'tensor_pipe_pct2' in rec:
            rec['tensor_pipe_pct'] = rec['tensor_pipe_pct2']
        rec.pop('tensor_pipe_pct2', None)
        launches.append(rec)
        g = per.setdefault(rec['kernel'], {'launches': 0, 'dram_bytes': 0.0, 'time_us': 0.0})
        g['launches'] += 1
        g['dram_bytes'] += rec['dram_bytes']
        g['time_us'] += rec.get('time_us', 0.0)
    for g in per.values():
        g['dram_bytes_per_launch'] = g['dram_bytes'] / g['launches']
    out = {'source': 'ncu --set full --clock-control none: ' + a.path, 'note': a.note, 'kernels': per, 'launches': launches}
    js = json.dumps(out, indent=1)
    if a.out:
        open(a.out + '_traffic.json', 'w').write(js)
    else:
        print(js)


def cmd_table(a):
    """per kernel of a *_traffic.json: launches, time, DRAM bytes, achieved GB/s and its fraction of the measured HBM peak"""
    d = json.load(open(a.path))
    agg = {}
    for l in d['launches']:
        g = agg.setdefault(l['kernel'], [0, 0.0, 0.0])
        g[0] += 1
        g[1] += l.get('time_us', 0.0)
        g[2] += l['dram_bytes']
    print('%-44s %4s %10s %10s %9s %6s' % ('kernel', 'n', 'time us', 'DRAM MB', 'GB/s', 'of HBM'))
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = by / us / 1e3 if us else 0.0
        print('%-44s %4d %10.1f %10.0f %9.0f %6.2f' % (k[:44], n, us, by / 1e6, gbs, gbs / a.peak))


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    l = sub.add_parser('launches')
    l.add_argument('path')
    l.add_argument('--first', type=int, default=None)
    l.add_argument('--count', type=int, default=None)
    f = sub.add_parser('full')
    f.add_argument('path')
    f.add_argument('--out', default=None)
    f.add_argument('--note', default='mean over the launches captured')
    t = sub.add_parser('table')
    t.add_argument('path')
    t.add_argument('--peak', type=float, default=6455.9, help='HBM GB/s (MEASURED_PEAKS.json)')
    a = ap.parse_args()
    {'launches': cmd_launches, 'full': cmd_full, 'table': cmd_table}[a.cmd](a)


if __name__ == '__main__':
    main()
