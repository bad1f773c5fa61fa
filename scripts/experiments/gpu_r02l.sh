#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest_gpu.log
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1; head -20 gpurun_out/${tag}_timeline.txt
MGB_SINE_MODES=0 python scripts/solve_timeline.py cfg5 2>&1 | head -3
timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['parity']['ok'], d['config']['conv'])
for k in d['kernels']: print('%-45s %.3f ms  hbm %.2f fp64 %.2f  x%d/it x%d/solve' % (k['name'],k['ms'],k.get('hbm_frac',0),k.get('fp64_frac',0),k['launches_per_iteration'],k['launches_per_solve']))
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --no-cpu --workload cfg2 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench cfg2 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_cfg2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['parity']['ok'], d['config']['conv'])
PY
