"""Tables of DESIGN.md section 7 from the bench lines under profiles/ (python scripts/design_table.py)."""
import glob
import json
import os

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r03'


def load(name):
    path = os.path.join(P, name)
    return json.load(open(path)) if os.path.exists(path) else None


print('| config (BASELINE.json) | hierarchy | iterations to 1e-10 | device, ms | **end to end, ms** (cold) | DOF/s end to end | launches | '
      'CPU: reference algorithm on the box (cores) | ratio e2e / CPU |')
print('|---|---|---|---|---|---|---|---|---|')
for wl in ('cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'):
    d = load(f'{TAG}_bench_{wl}_n1.json')
    if d is None:
        continue
    w = d['config']['workload']
    cpu = d.get('cpu_baseline') or {}
    ratio = d['e2e']['value'] / cpu['value'] if cpu.get('value') else float('nan')
    import re
    what = re.match(r'[^:]+: (.*?) on \[', w).group(1)
    hier = re.search(r'(\d-level, coarsening [0-9x]+)', w).group(1)
    print(f"| {wl}: {what} | {hier} | {d['config']['iterations']} | {d['ms_per_step']:.2f} | "
          f"**{d['e2e']['ms_per_step']:.2f}** ({d['e2e']['cold_ms']:.0f}) | {d['e2e']['value']:.2e} | {d['gpu_launches']} | "
          f"{cpu.get('value', float('nan')):.2e} ({cpu.get('cores')}, {cpu.get('kind')}) | {ratio:.0f}x |")
print()
print('| cfg 5 on N GPUs | 1 | 2 | 4 | 8 |')
print('|---|---|---|---|---|')
rows = {n: load(f'{TAG}_bench_cfg5_n{n}.json') for n in (1, 2, 4, 8)}
def cell(fn):
    return ' | '.join(fn(rows[n]) if rows[n] else '-' for n in (1, 2, 4, 8))
print('| device time to 1e-10, ms | ' + cell(lambda d: f"{d['ms_per_step']:.2f}") + ' |')
print('| end to end, ms | ' + cell(lambda d: f"{d['e2e']['ms_per_step']:.2f}") + ' |')
print('| space-time DOF/s (device) | ' + cell(lambda d: f"{d['value']:.2e}") + ' |')
print('| launches per solve and rank | ' + cell(lambda d: str(d['gpu_launches'])) + ' |')
print('| parity block: points checked, max rel. error at F-points, history vs N=1 | ' +
      cell(lambda d: f"{d['parity']['points']}, {d['parity']['max_rel']:.1e}, {d['parity'].get('conv_abs_diff', 0):.1e}") + ' |')
print()
d = rows[1]
print('| level-0 sweep of cfg 5 (16 384 intervals of 64) | ms | GB/s | of measured HBM peak | of FP64 issue | per iteration / per solve |')
print('|---|---|---|---|---|---|')
for k in d['kernels']:
    print(f"| {k['name']} | {k['ms']:.3f} | {k['gbs']:.0f} | {k.get('hbm_frac', 0):.2f} | {k.get('fp64_frac', 0):.2f} | "
          f"{k['launches_per_iteration']} / {k['launches_per_solve']} |")

for wl, title in (('cfg3', 'level-0 sweep of cfg 3 (512 intervals of 8, rows of 2 MiB)'),
                  ('cfg4', 'level-0 sweep of cfg 4 (32 768 intervals of 2, rows of 32 KiB)')):
    d = load(f'{TAG}_bench_{wl}_n1.json')
    if d is None:
        continue
    print()
    print(f'| {title} | ms | GB/s | of measured HBM peak | per iteration / per solve |')
    print('|---|---|---|---|---|')
    for k in d['kernels']:
        print(f"| {k['name']} | {k['ms']:.3f} | {k['gbs']:.0f} | {k.get('hbm_frac', 0):.2f} | "
              f"{k['launches_per_iteration']} / {k['launches_per_solve']} |")
