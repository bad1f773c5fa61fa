"""Coarsest-level solve of the headline hierarchy alone: sine-space solve vs the sequential Phi chain.

    python scripts/bench_forward.py [npts]
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 513
kw = {k: v for k, v in bench.HEAT_KW.items() if k not in ('t_start', 't_stop')}
fine = P.Heat1D(nt=2 * (npts - 1) + 1, **bench.HEAT_KW)
coarse = P.Heat1D(t_interval=fine.t[::2], **kw)
solver = P.Mgrit(problem=[fine, coarse], nested_iteration=False, logging_lvl=logging.WARNING, tol=1e-10)
lv = solver._lv[1]
lv.g[:, :lv.n].normal_()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


sp = solver._spectral.get(1)
print(f'{npts} points: sine-space solve {timed(lambda: solver.forward_solve(1)):.1f} us '
      f'(transform in {timed(sp.transform_in):.1f}, recurrences {timed(sp.recur):.1f}, transform out {timed(sp.transform_out):.1f})')
solver._spectral.clear()
print(f'{npts} points: Phi chain {timed(lambda: solver.forward_solve(1)):.1f} us')
