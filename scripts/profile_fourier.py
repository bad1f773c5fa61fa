"""Launch the coarsest-level solve of cfg4 (advection nx = 4096, nt = 65537, two levels: 32769 coarse points) in Fourier
space twice: a warm-up pass, then the pass ncu captures.

    ncu --set full --clock-control none --import-source on -k "regex:k_rows_rfft|k_rows_irfft|k_cplx_solve" -s 3 -c 3 \
        -o gpurun_out/prof_fourier python scripts/profile_fourier.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

w = bench.WORKLOADS['cfg4']
prob = bench.build_levels(P.Advection1D, w['kw'], w['t'], w['coarsening'])
solver = P.Mgrit(problem=prob, logging_lvl=logging.WARNING, **dict(w['solver'], nested_iteration=False))
last = len(prob) - 1
for _ in range(2):
    solver.forward_solve(last)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
solver.forward_solve(last)
e1.record()
torch.cuda.synchronize()
print('coarsest solve of %d points: %.3f ms' % (solver._lv[last].npts, e0.elapsed_time(e1)))
