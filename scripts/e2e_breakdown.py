"""Where the end-to-end time of the headline workload goes (host setup vs device), run on the GPU box:

    python scripts/e2e_breakdown.py [cfg5|cfg2]
"""
import cProfile
import logging
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

import torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:                       # python -m torch.distributed.run --nproc-per-node N scripts/e2e_breakdown.py
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    if rank != 0:
        sys.stdout = open(os.devnull, 'w')
wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
nt, co = bench.workload_grid(wl)


def run():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    prob = bench.hierarchy(P.Heat1D, nt, co)
    t.append(time.perf_counter())
    s = P.Mgrit(problem=prob, logging_lvl=logging.WARNING, **bench.SOLVER_KW)
    t.append(time.perf_counter())
    info = s.solve()
    t.append(time.perf_counter())
    last = s.u[0][-1].get_values()
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])], s


for k in range(8):
    ms, s = run()
    print('hierarchy %.1f ms | Mgrit() %.1f ms | solve() %.1f ms | readback %.1f ms | total %.1f ms' % (*ms, sum(ms)))
    print('    ' + ' | '.join('%s %.2f' % ph for ph in s.setup_phases))
    del s
# finer marks of one more run (core/device_level.py mark())
from pymgrit_b200.core import device_level as dl
dl.TRACE = []
t_begin = time.perf_counter()
ms, s = run()
marks, dl.TRACE = dl.TRACE, None
prev = t_begin
print('marks of one run (ms since the previous mark):')
for label, tm in marks:
    print('    %-36s %7.3f   (at %7.3f)' % (label, 1e3 * (tm - prev), 1e3 * (tm - t_begin)))
    prev = tm
del s
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(60)
pstats.Stats(pr).sort_stats('tottime').print_stats(40)
