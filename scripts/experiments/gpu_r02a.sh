#!/bin/bash
# r02 first visit: parity tests, bench line without the CPU leg, timeline, e2e breakdown
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-2500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1; cat gpurun_out/${tag}_timeline.txt
MGB_HEAT1D_SINE=0 python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline_node.txt 2>&1; cat gpurun_out/${tag}_timeline_node.txt
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown.txt 2>&1
head -8 gpurun_out/${tag}_e2e_breakdown.txt
