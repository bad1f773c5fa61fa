"""Device-time breakdown of one solve of the headline workload: CUDA events around every sweep, summed per (sweep, level).

    python scripts/solve_timeline.py [cfg1..cfg5] [coarsening, e.g. 16,16,8]

Also prints host-side wall time of restart()+solve() so the gap between the device busy time and the wall time
(launch overhead, synchronisations in the convergence test) is visible.
"""
import collections
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P
from pymgrit_b200.core import mgrit as M

import torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:                       # python -m torch.distributed.run --nproc-per-node N scripts/solve_timeline.py
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
show = int(os.environ.get('SHOW_RANK', world - 1))
if rank != show:
    sys.stdout = open(os.devnull, 'w')
wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
nt, co = bench.workload_grid(wl)
if len(sys.argv) > 2:
    co = tuple(int(x) for x in sys.argv[2].split(','))

w = bench.WORKLOADS[wl]
solver = P.Mgrit(problem=bench.build_levels(bench.app_class(P, w['app']), w['kw'], w['t'], co), logging_lvl=logging.WARNING,
                 **w['solver'])
for _ in range(2):
    solver.restart()
    solver.solve()

# plain wall / device time of restart + solve
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
solver.restart()
info = solver.solve()
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
print(f'{bench.describe(wl, co)}: {len(info["conv"])} iterations, device {e0.elapsed_time(e1):.2f} ms, wall {wall:.2f} ms')

records = []


def wrap(name):
    fn = getattr(M.Mgrit, name)

    def inner(self, *a, **kw):
        lvl = kw.get('lvl', a[0] if a else 0)
        tag = name + ('(last)' if kw.get('last_only') else '')
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn(self, *a, **kw)
        e.record()
        records.append((tag, lvl, s, e))
        return r
    setattr(M.Mgrit, name, inner)


for nm in ('f_relax', 'c_relax', 'fas_residual', 'down_sweep', 'error_correction', 'forward_solve', 'convergence_criterion',
           '_exchange_ghost'):
    wrap(nm)
_orig_nested = M.Mgrit.nested_iteration
solver.restart()
n_setup = len(records)
solver.solve()
torch.cuda.synchronize()
tot = collections.OrderedDict()
for k, (tag, lvl, s, e) in enumerate(records):
    key = ('setup ' if k < n_setup else 'iter  ') + f'{tag:28s} lvl {lvl}'
    a = tot.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += s.elapsed_time(e)
total = sum(v[1] for v in tot.values())
for key, (cnt, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{key}  x{cnt:3d}  {ms:8.3f} ms  {100 * ms / total:5.1f}%')
print(f'sum of sweeps {total:.2f} ms (events add their own gaps; _exchange_ghost is nested inside c_relax)')
if world > 1:
    dist.destroy_process_group()
