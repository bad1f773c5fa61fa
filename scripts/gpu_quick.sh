#!/bin/bash
# Short GPU-box visit: parity tests + bench line (+ e2e breakdown).  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag>'
tag=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json | cut -c1-1500
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown.txt 2>&1
head -5 gpurun_out/${tag}_e2e_breakdown.txt
