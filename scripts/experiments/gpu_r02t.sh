#!/bin/bash
# FFT v3 + helper-thread level-0 tables: selected tests, cfg4 timeline, e2e A/B, cfg5 bench.
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rfft or fourier or advection or Advection or cfg4 or heat1d or Heat1D or example" > gpurun_out/${tag}_pytest_sel.log 2>&1; echo "pytest sel rc=$?"
tail -6 gpurun_out/${tag}_pytest_sel.log
echo "== cfg4 timeline (2 levels)"
timeout 300 python scripts/solve_timeline.py cfg4 2>&1 | head -10
timeout 300 python scripts/profile_fourier.py
echo "== e2e breakdown, tables ahead"
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
sed -n '1,16p' gpurun_out/${tag}_e2e_breakdown_n1.txt | cut -c1-260
sed -n '/marks of one run/,/function calls/p' gpurun_out/${tag}_e2e_breakdown_n1.txt | head -50
echo "== e2e breakdown, MGB_TABLES_AHEAD=0"
MGB_TABLES_AHEAD=0 timeout 300 python scripts/e2e_breakdown.py cfg5 2>&1 | sed -n '9,16p' | cut -c1-260
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02t_bench_cfg5.json'.replace('r02t', __import__('os').environ.get('TAG','r02t'))))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('cold_ms'), d['gpu_launches'], d['parity']['ok'])
PY
