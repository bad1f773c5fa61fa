"""Summaries of ncu outputs for profiles/ (run in the build container, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv      # per-kernel totals of a launch list
    python scripts/ncu_summary.py full gpurun_out/r01_sweeps.ncu-rep        # key metrics of a --set full capture
    python scripts/ncu_summary.py traffic gpurun_out/r01_sweeps.ncu-rep name1,name2,... <intervals> profiles/ncu_traffic.json [m]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name[:110]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        if r[mi] != 'gpu__time_duration.sum':
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
        k = short(r[ki])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    total = sum(tot.values())
    print(f'# {path}: {sum(cnt.values())} launches, {total / 1e3:.3f} ms of kernel time (ncu: cold cache, serialised)')
    print(f'{"share":>7} {"total_us":>11} {"count":>6} {"avg_us":>10}  kernel')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f'{100 * v / total:6.2f}% {v:11.1f} {cnt[k]:6d} {v / cnt[k]:10.2f}  {k}')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum']


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index('Kernel Name')
    idx = [(w, h.index(w)) for w in WANT if w in h]
    print(f'# {path}: {len(rows) - 2} captured launches (ncu --set full --clock-control none)')
    for r in rows[2:]:
        print(short(r[ki]))
        for w, i in idx:
            print(f'    {w:62s} {r[i]:>16s} {units[i]}')
        try:
            rd = float(r[h.index('dram__bytes_read.sum')].replace(',', ''))
            wr = float(r[h.index('dram__bytes_write.sum')].replace(',', ''))
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            rd *= scale[units[h.index('dram__bytes_read.sum')]]
            wr *= scale[units[h.index('dram__bytes_write.sum')]]
            us = float(r[h.index('gpu__time_duration.sum')].replace(',', ''))
            us *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[h.index('gpu__time_duration.sum')]]
            print(f'    {"-> dram traffic (read+write)":62s} {(rd + wr) / 1e9:16.4f} GB  = {(rd + wr) / us / 1e3:.0f} GB/s under ncu')
        except Exception:
            pass


def traffic(path, names, intervals, out_path, coarsening=None):
    """profiles/ncu_traffic.json for bench.py: DRAM bytes (read + write) per captured launch, keyed by the sweep names
    given in launch order (scripts/profile_sweeps.py), with the number of coarse intervals the capture ran on."""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    ir, iw = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    res = {'source': path, 'intervals': int(intervals), 'sweeps': {}}
    if coarsening is not None:
        res['coarsening'] = int(coarsening)          # length of the level-0 intervals of the capture (bench.py checks it)
    for name, r in zip(names.split(','), rows[2:]):
        rd = float(r[ir].replace(',', '')) * scale[units[ir]]
        wr = float(r[iw].replace(',', '')) * scale[units[iw]]
        res['sweeps'][name] = {'kernel': short(r[h.index('Kernel Name')]), 'dram_read_bytes': rd, 'dram_write_bytes': wr}
    with open(out_path, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    if sys.argv[1] == 'traffic':
        traffic(*sys.argv[2:7])
    else:
        {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
