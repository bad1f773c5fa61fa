"""Launch every level-0 sweep of the headline workload twice (warm-up pass, then the pass ncu captures).

    ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_c_relax|k_fas_residual|k_down|k_correct|k_residual" -s 7 -c 7 -o gpurun_out/prof python scripts/profile_sweeps.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

nt = int(os.environ.get('NT', 2 ** 20 + 1))
levels = int(os.environ.get('LEVELS', 8))
m = int(os.environ.get('COARSENING', 4))
problem = P.simple_setup_problem(P.Heat1D(nt=nt, **bench.HEAT_KW), level=levels, coarsening=m)
solver = P.Mgrit(problem=problem, nested_iteration=False, logging_lvl=logging.WARNING, tol=1e-10)
for _ in range(2):
    solver.f_relax(0)
    solver.f_relax(0, last_only=True)
    solver.c_relax(0)
    solver.fas_residual(0)
    solver.down_sweep(0)
    solver.error_correction(0, f_relax=True)
    solver.compute_residual()
    torch.cuda.synchronize()
print('done')
