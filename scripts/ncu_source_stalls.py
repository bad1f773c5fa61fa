"""Aggregate an `ncu --page source --csv` export: stall reasons and instruction mix per kernel launch.

    ncu -i rep.ncu-rep --page source --csv > src.csv; python scripts/ncu_source_stalls.py src.csv [units_per_launch ...]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = [float(x) for x in sys.argv[2:]]
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        sections.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for k, sec in enumerate(sections):
    h, data = sec['rows'][0], [r for r in sec['rows'][1:] if len(r) == len(sec['rows'][0])]
    iS, isrc, iE = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    agg = collections.Counter()
    for r in data:
        for i in stall:
            agg[h[i]] += int(r[i])
    tot = max(sum(agg.values()), 1)
    ex = sum(int(r[iE]) for r in data)
    u = units[k] if k < len(units) else 1.0
    print(f'== launch {k}: {sec["name"][:90]}')
    print('   stalls %:', ', '.join(f'{n[6:]} {100 * v / tot:.1f}' for n, v in agg.most_common(8)))
    print(f'   warp instructions executed: {ex} ({ex / u:.1f} per unit)')
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        s = r[isrc].split()
        o = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
        op[o] += int(r[iE])
        ops[o] += int(r[iS])
    ts = max(sum(ops.values()), 1)
    print('   opcode: executed per unit / % of samples:',
          ', '.join(f'{o} {c / u:.1f}/{100 * ops[o] / ts:.1f}%' for o, c in op.most_common(14)))
