"""Launch the level-0 sweeps of cfg3 (heat_2d 512 x 512, nt = 4097, coarsening 8) and the node <-> sine-space transforms
twice: a warm-up pass, then the pass ncu captures.

    ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual|k_dgemm" -s 7 -c 7 \
        -o gpurun_out/prof_cfg3 python scripts/profile_cfg3.py
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymgrit_b200 as P

w = bench.WORKLOADS['cfg3']
prob = bench.build_levels(P.Heat2D, w['kw'], w['t'], w['coarsening'])
solver = P.Mgrit(problem=prob, logging_lvl=logging.WARNING, nested_iteration=False, **{k: v for k, v in w['solver'].items()})
for _ in range(2):
    solver.f_relax(0, last_only=True)                               # k_chain
    solver.down_sweep(0)                                            # k_down
    solver.error_correction(0, f_relax=True, last_only=True)        # k_correct
    solver.compute_residual()                                       # k_residual (+ k_sum_systems)
    solver.f_relax(0)                                               # k_chain (every F-point stored)
    vals = solver.u[0][7].get_values()                              # 2 x k_dgemm: sine space -> node values
    torch.cuda.synchronize()
print('done', vals.shape)
