"""cfg4 (advection nx=4096, nt=65537, coarsening 2, nested iteration): iterations and device time to 1e-10 for several
depths of the hierarchy and both cycle types (BASELINE.json leaves the number of levels to the builder).

    python scripts/cfg4_sweep.py
"""
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

w = bench.WORKLOADS['cfg4']
for levels in (2, 3, 4, 5, 7, 10):
    for cyc in ('V', 'F'):
        if levels == 2 and cyc == 'F':
            continue
        prob = bench.build_levels(P.Advection1D, w['kw'], w['t'], (2,) * (levels - 1))
        kw = dict(w['solver'], cycle_type=cyc, max_iter=300)
        s = P.Mgrit(problem=prob, logging_lvl=logging.WARNING, **kw)
        s.solve()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.restart()
        info = s.solve()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        c = info['conv']
        print(f'levels {levels:2d} {cyc}-cycle: {len(c):3d} iterations, last conv {c[-1]:.2e}, {ms:8.1f} ms', flush=True)
        del s
