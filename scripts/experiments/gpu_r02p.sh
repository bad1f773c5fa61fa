#!/bin/bash
# Heat2D on the one-thread-per-mode sweeps: parity (team vs modes, fixtures, full size), cfg3 bench line and timeline,
# grid-size sweep.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r02p.sh <tag>'
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "heat2d or Heat2D or cfg3 or one_thread" > gpurun_out/${tag}_pytest_heat2d.log 2>&1; echo "pytest heat2d rc=$?"
tail -8 gpurun_out/${tag}_pytest_heat2d.log
for cap in 64 16 8 4; do
  echo "== MGB_MODES_CTAS_PER_SM=$cap"
  MGB_MODES_CTAS_PER_SM=$cap timeout 300 python scripts/solve_timeline.py cfg3 2>&1 | head -24
done
echo "== team kernels"
MGB_SINE_MODES=0 timeout 300 python scripts/solve_timeline.py cfg3 2>&1 | head -12
timeout 600 python bench.py --workload cfg3 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/${tag}_bench_cfg3.json
