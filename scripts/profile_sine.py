"""Launch the level-0 sweeps of the headline workload (sine-space rows unless MGB_HEAT1D_SINE=0) twice: a warm-up pass,
then the pass ncu captures.

    ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual" -s 5 -c 5 \
        -o gpurun_out/prof python scripts/profile_sine.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

nt, co = bench.workload_grid('cfg5')
solver = P.Mgrit(problem=bench.hierarchy(P.Heat1D, nt, co), logging_lvl=logging.WARNING, nested_iteration=False, tol=1e-10)
for _ in range(2):
    solver.f_relax(0, last_only=True)                               # k_chain
    solver.down_sweep(0)                                            # k_down
    solver.error_correction(0, f_relax=True, last_only=True)        # k_correct
    solver.compute_residual()                                       # k_residual
    solver.f_relax(0)                                               # k_chain (every F-point stored)
    torch.cuda.synchronize()
print('done')
