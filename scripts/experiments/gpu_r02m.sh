#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1; head -14 gpurun_out/${tag}_timeline.txt
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
head -12 gpurun_out/${tag}_e2e_breakdown_n1.txt
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "mode or lazy or fused or sine" 2>&1 | tail -3
