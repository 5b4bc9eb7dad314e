#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one row per profiled launch with the counters the roofline discussion uses.
usage: ncu_summary.py report.ncu-rep out_prefix     -> out_prefix.csv (table) and out_prefix_traffic.json (DRAM bytes per kernel)"""
import csv
import io
import json
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('lts__t_bytes.sum', 'l2_bytes'), ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts'),
        ('sm__cycles_elapsed.avg', 'cycles')]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    table, traffic = [], {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('<unnamed>::', '')
        rec = {'id': r[idx['ID']], 'kernel': name}
        for m, short in WANT:
            if m not in idx:
                continue
            v = r[idx[m]].replace(',', '')
            try:
                v = float(v) * UNIT.get(units[idx[m]], 1.0)
            except ValueError:
                pass
            rec[short] = v
        table.append(rec)
        t = traffic.setdefault(name, {'launches': 0, 'dram_bytes': 0.0, 'time_us': 0.0})
        t['launches'] += 1
        t['dram_bytes'] += rec.get('dram_rd', 0.0) + rec.get('dram_wr', 0.0)
        t['time_us'] += rec.get('time', 0.0)
    cols = ['id', 'kernel'] + [s for _, s in WANT]
    with open(out + '.csv', 'w') as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for rec in table:
            w.writerow({c: rec.get(c, '') for c in cols})
    for t in traffic.values():
        t['dram_bytes_per_launch'] = t['dram_bytes'] / t['launches']
    with open(out + '_traffic.json', 'w') as f:
        json.dump({'source': rep.split('/')[-1], 'units': {'time': 'us', 'bytes': 'B'}, 'kernels': traffic,
                   'launches': [{'id': r['id'], 'kernel': r['kernel'], 'time_us': r.get('time'), 'dram_bytes': r.get('dram_rd', 0) + r.get('dram_wr', 0)}
                                for r in table]}, f, indent=1)
    for rec in table:
        print(' '.join('%s=%s' % (c, ('%.4g' % rec[c]) if isinstance(rec.get(c), float) else rec.get(c)) for c in cols))


if __name__ == '__main__':
    main()
