#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['parity']['ok'], d['config']['conv'])
for k in d['kernels']: print('%-45s %.3f ms  hbm %.2f fp64 %.2f  x%d/it x%d/solve' % (k['name'],k['ms'],k.get('hbm_frac',0),k.get('fp64_frac',0),k['launches_per_iteration'],k['launches_per_solve']))
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['parity']['ok'], d['config']['conv'])
PY
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
head -16 gpurun_out/${tag}_e2e_breakdown_n1.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n2.txt 2>&1
grep -A1 "^hierarchy" gpurun_out/${tag}_e2e_breakdown_n2.txt | head -16
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1; head -20 gpurun_out/${tag}_timeline.txt
for wl in cfg1 cfg2 cfg3 cfg4; do
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
