#!/bin/bash
# Radix-8 two-rows-per-transform FFT + fused norm/flag: full GPU test suite, cfg4 sweep, timelines, cfg1/cfg2/cfg5 bench.
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${tag}_pytest_gpu.log
echo "== cfg4 sweep (Fourier coarsest solve)"
timeout 600 python scripts/cfg4_sweep.py 2>&1 | tee gpurun_out/${tag}_cfg4_sweep.txt
echo "== cfg4 timeline"
timeout 300 python scripts/solve_timeline.py cfg4 2>&1 | head -12
echo "== cfg4 timeline 2 levels"
timeout 300 python scripts/solve_timeline.py cfg4 2 2>&1 | head -12
echo "== cfg5 timeline"
timeout 300 python scripts/solve_timeline.py cfg5 2>&1 | head -8
echo "== cfg2 timeline"
timeout 300 python scripts/solve_timeline.py cfg2 2>&1 | head -4
