#!/bin/bash
# predicted-last-cycle full store: tests, cfg5 / cfg3 / cfg4 bench with and without (MGB_PREDICT_LAST=0)
tag=${1:-r03m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
for p in 1 0; do
  MGB_PREDICT_LAST=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_cfg5_p$p.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_cfg5_p$p.json'))
print('predict=$p cfg5', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'], d['parity']['ok'], d['config']['iterations'], [(k['name'][:28], k.get('launches_last_solve')) for k in d['kernels'] if k.get('launches_last_solve')], d['roofline']['kernel'], round(d['roofline']['frac'],3))
PY
done
for wl in cfg3 cfg4; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_${wl}.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_${wl}.json'))
print('$wl', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'], d['parity']['ok'], d['config']['iterations'])
PY
done
