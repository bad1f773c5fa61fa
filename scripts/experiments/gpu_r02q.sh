#!/bin/bash
# Advection coarsest solve in Fourier space + blocked down-sweep: tests, cfg4 hierarchy sweep, timelines.
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rfft or fourier or advection or Advection or heat2d or one_thread or cfg4 or cfg3" > gpurun_out/${tag}_pytest_sel.log 2>&1; echo "pytest sel rc=$?"
tail -15 gpurun_out/${tag}_pytest_sel.log
echo "== cfg4 sweep (Fourier coarsest solve)"
timeout 600 python scripts/cfg4_sweep.py 2>&1 | tee gpurun_out/${tag}_cfg4_sweep.txt
echo "== cfg4 timeline"
timeout 300 python scripts/solve_timeline.py cfg4 2>&1 | head -30
echo "== cfg3 timeline"
timeout 300 python scripts/solve_timeline.py cfg3 2>&1 | head -14
echo "== cfg5 timeline"
timeout 300 python scripts/solve_timeline.py cfg5 2>&1 | head -14
