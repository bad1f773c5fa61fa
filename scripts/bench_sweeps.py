"""CUDA-event timing of the level-0 sweeps of the headline workload (quick A/B runs on the GPU box).

    [NT=..] [COARSENING=16] python scripts/bench_sweeps.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

nt = int(os.environ.get('NT', 2 ** 20 + 1))
m = int(os.environ.get('COARSENING', 16))
kw = dict(bench.HEAT_KW)
if os.environ.get('RHS', '1') == '0':
    kw.pop('rhs')
problem = P.simple_setup_problem(P.Heat1D(nt=nt, **kw), level=3, coarsening=m)
solver = P.Mgrit(problem=problem, nested_iteration=False, logging_lvl=logging.WARNING, tol=1e-10)
for k in solver.time_level0_sweeps(repeats=10):
    print('%-28s %8.3f ms  %7.0f GB/s' % (k['name'], k['ms'], k['gbs']))
# sequential solve on a coarse grid of 513 points (the coarsest level of the bench hierarchy)
import numpy as np
coarse = P.Heat1D(t_interval=np.linspace(0, 2, 513), **{k: v for k, v in bench.HEAT_KW.items() if k not in ('t_start', 't_stop')})
s1 = P.Mgrit(problem=[coarse], logging_lvl=logging.WARNING, nested_iteration=False)
s1.forward_solve(0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    s1.forward_solve(0)
e1.record()
torch.cuda.synchronize()
print('forward_solve 513 points      %8.3f ms  (%.2f us per step)' % (e0.elapsed_time(e1) / 10, e0.elapsed_time(e1) / 10 / 512 * 1e3))
