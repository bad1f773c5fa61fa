"""Opcode histogram per kernel of libmgrit_b200.so (cuobjdump -sass), so that what the kernels are made of can be checked
from the tree: bulk (TMA) copies UBLKCP, mbarrier waits SYNCS, FP64 arithmetic DFMA/DMUL/DADD, plain global loads and
stores LDG/STG, shuffles, local-memory spills LDL/STL.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pymgrit_b200', 'lib', 'libmgrit_b200.so')
WANT = ['UBLKCP', 'SYNCS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'LDL', 'STL', 'BAR', 'ATOM', 'RED']

out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
kern, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*arch = (\S+)', line)
    if m:
        arch.add(m.group(1))
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and kern:
        counts[kern][m.group(1)] += 1
        counts[kern]['_all'] += 1


def demangle(names):
    r = subprocess.run(['c++filt'], input='\n'.join(names), stdout=subprocess.PIPE, text=True)
    return r.stdout.splitlines()


names = list(counts)
pretty = demangle(names)
print(f'# {os.path.relpath(LIB, ROOT)}: {len(names)} kernels, arch {sorted(arch)}; static SASS instruction counts per kernel')
print('# ' + ' '.join(f'{w:>7}' for w in ['all'] + WANT) + '  kernel')
tot = collections.Counter()
rows = []
for n, p in zip(names, pretty):
    c = counts[n]
    p = re.sub(r'^void ', '', p)
    p = re.sub(r'\(.*$', '', p)
    p = p.replace('mgb::', '').replace('(int)', '')
    rows.append((p, c))
    tot.update(c)
for p, c in sorted(rows, key=lambda r: r[0]):
    print('  ' + ' '.join(f'{c.get(w, 0):7d}' for w in ['_all'] + WANT) + '  ' + p[:100])
print('  ' + ' '.join(f'{tot.get(w, 0):7d}' for w in ['_all'] + WANT) + '  TOTAL')
