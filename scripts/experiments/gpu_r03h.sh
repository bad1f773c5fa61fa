#!/bin/bash
tag=${1:-r03}
mkdir -p gpurun_out
for wl in cfg1 cfg2 cfg3; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${wl}_n1.json 2> gpurun_out/${tag}_bench_${wl}_n1.err; echo "bench $wl rc=$?"
done
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg4_n1.json 2> gpurun_out/${tag}_bench_cfg4_n1.err; echo "bench cfg4 rc=$?"
python - <<PY
import json
for wl in ('cfg1','cfg2','cfg3','cfg4'):
    d=json.load(open('gpurun_out/${tag}_bench_%s_n1.json' % wl))
    print(wl, 'device %.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d['parity'] and d['parity']['ok'])
PY
