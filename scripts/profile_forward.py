"""The sequential coarsest-level solve alone (latency-bound): time per step, and a target for ncu's source view.

    python scripts/profile_forward.py [npts]
    ncu --set full --import-source on --clock-control none -k regex:k_chain -s 3 -c 1 -o gpurun_out/fwd \
        python scripts/profile_forward.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 513
kw = {k: v for k, v in bench.HEAT_KW.items()}
fine = P.Heat1D(nt=2 * (npts - 1) + 1, **kw)
coarse = P.Heat1D(t_interval=fine.t[::2], **{k: v for k, v in kw.items() if k not in ('t_start', 't_stop')})
solver = P.Mgrit(problem=[fine, coarse], nested_iteration=False, logging_lvl=logging.WARNING, tol=1e-10)
solver._lv[1].g.normal_()
for _ in range(3):
    solver.forward_solve(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    solver.forward_solve(1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'forward_solve over {npts} points with g: {ms * 1e3:.1f} us = {ms * 1e3 / (npts - 1):.3f} us per step')
