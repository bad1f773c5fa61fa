#!/bin/bash
tag=${1:-r03g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
sed -n '9,16p' gpurun_out/${tag}_e2e_breakdown_n1.txt | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n1.json 2> gpurun_out/${tag}_bench_cfg5_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_cfg5_n1.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['gpu_launches'], d['parity']['ok'])
PY
for wl in cfg2 cfg3; do timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_${wl}_n1.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_${wl}_n1.json'))
print('$wl', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['gpu_launches'], d['parity']['ok'])
PY
done
