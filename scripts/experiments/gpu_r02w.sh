#!/bin/bash
tag=${1:-r02w}
mkdir -p gpurun_out
python scripts/solve_timeline.py cfg5 2>&1 | head -14
python scripts/solve_timeline.py cfg2 2>&1 | head -3
python scripts/solve_timeline.py cfg3 2>&1 | head -8
python - <<'PY'
import logging, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
import pymgrit_b200 as P
w = bench.WORKLOADS['cfg5']
s = P.Mgrit(problem=bench.build_levels(P.Heat1D, w['kw'], w['t'], w['coarsening']), logging_lvl=logging.WARNING, **w['solver'])
s.solve()
for k in s.time_level0_sweeps(repeats=10, iterations=3, hbm_gbs=6550.4):
    print('%-45s %.4f ms' % (k['name'], k['ms']))
PY
