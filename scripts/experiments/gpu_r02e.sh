#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
for shp in 64,17 32,33; do
  echo "== sine shape $shp"
  MGB_SHAPE_7=$shp python scripts/solve_timeline.py cfg5 2>&1 | head -14
done > gpurun_out/${tag}_sine_shapes.txt 2>&1
cat gpurun_out/${tag}_sine_shapes.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['parity'])
for k in d['kernels']: print('%-45s %.3f ms  hbm %.2f fp64 %.2f  x%d/it x%d/solve' % (k['name'],k['ms'],k.get('hbm_frac',0),k.get('fp64_frac',0),k['launches_per_iteration'],k['launches_per_solve']))
PY
tail -5 gpurun_out/${tag}_bench.err
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown.txt 2>&1
head -45 gpurun_out/${tag}_e2e_breakdown.txt
