#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1; head -20 gpurun_out/${tag}_timeline.txt
for wl in cfg5 cfg3; do
  timeout 900 python bench.py --workload $wl --no-cpu > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench_${wl}.json'))
    print('$wl', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['config']['iterations'], d['parity'] and d['parity']['ok'], d['config']['conv'][-1])
    for k in d['kernels']: print('   %-45s %.3f ms  hbm %.2f  x%d/it x%d/solve' % (k['name'],k['ms'],k.get('hbm_frac',0),k['launches_per_iteration'],k['launches_per_solve']))
except Exception as e: print('$wl failed', e)
PY
  tail -3 gpurun_out/${tag}_bench_${wl}.err
done
timeout 900 python scripts/cfg4_sweep.py > gpurun_out/${tag}_cfg4_sweep.txt 2>&1; cat gpurun_out/${tag}_cfg4_sweep.txt
