#!/bin/bash
# FFT v4 (squared twiddles, permuted chirp spectrum) + native host passes: tests, cfg4 timeline, e2e breakdown, bench cfg5.
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_pytest_gpu.log
echo "== cfg4 timeline (2 levels)"
timeout 300 python scripts/solve_timeline.py cfg4 2>&1 | head -6
timeout 300 python scripts/profile_fourier.py
echo "== e2e breakdown"
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
sed -n '9,16p' gpurun_out/${tag}_e2e_breakdown_n1.txt | cut -c1-260
sed -n '/marks of one run/,/function calls/p' gpurun_out/${tag}_e2e_breakdown_n1.txt | head -40
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_cfg5.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('cold_ms'), d['gpu_launches'], d['parity']['ok'])
PY
