"""Time-to-tolerance of the headline workload for several time-grid hierarchies (coarsening factors per level).

    python scripts/hierarchy_sweep.py            # prints one line per hierarchy
"""
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import pymgrit_b200 as P

import torch.distributed as dist
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:                       # python -m torch.distributed.run --nproc-per-node N scripts/hierarchy_sweep.py ...
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    if rank != 0:
        sys.stdout = open(os.devnull, 'w')
nt = int(os.environ.get('NT', 2 ** 20 + 1))
CONFIGS = [
    [4] * 7, [4] * 5, [8] * 4, [8] * 5, [16] * 3, [16] * 4, [32] * 3, [16, 4, 4, 4], [16, 8, 8], [32, 8, 4], [8, 4, 4, 4, 4],
    [64, 16, 16], [16, 16, 8],
]
if len(sys.argv) > 1:
    CONFIGS = [[int(x) for x in a.split(',')] for a in sys.argv[1:]]
for ms in CONFIGS:
    fine = P.Heat1D(nt=nt, **bench.HEAT_KW)
    problem = [fine]
    for m in ms:
        problem.append(P.Heat1D(t_interval=problem[-1].t[::m], **{k: v for k, v in bench.HEAT_KW.items()
                                                                    if k not in ('t_start', 't_stop')}))
    solver = P.Mgrit(problem=problem, logging_lvl=logging.WARNING, **bench.SOLVER_KW)
    info = solver.solve()
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.restart()
        info = solver.solve()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    print(f'ranks={world} m={ms} coarsest={len(problem[-1].t)} iters={len(info["conv"])} conv={info["conv"][-1]:.2e} '
          f'ms={1e3 * min(times):.2f}', flush=True)
    del solver, problem
    torch.cuda.empty_cache()
