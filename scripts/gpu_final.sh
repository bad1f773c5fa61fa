#!/bin/bash
# Final single-GPU pass of the round: GPU tests, smoke(), the bench line of every config (with the CPU baseline), the
# reference arm.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_final.sh r03'
tag=${1:-r03}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n1.json 2> gpurun_out/${tag}_bench_cfg5_n1.err; echo "bench cfg5 rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench_cfg5_n1.err
for wl in cfg1 cfg2 cfg3; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${wl}_n1.json 2> gpurun_out/${tag}_bench_${wl}_n1.err; echo "bench $wl rc=$?"
done
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg4_n1.json 2> gpurun_out/${tag}_bench_cfg4_n1.err; echo "bench cfg4 rc=$?"
python - <<PY
import json
for wl in ('cfg5','cfg1','cfg2','cfg3','cfg4'):
    d=json.load(open('gpurun_out/${tag}_bench_%s_n1.json' % wl))
    print(wl, 'device %.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], 'cold %.0f' % d['e2e']['cold_ms'], 'launches', d['gpu_launches'], 'parity', d['parity'] and d['parity']['ok'], 'roofline', round(d['roofline']['frac'], 3) if d.get('roofline') else None, d['roofline']['kernel'] if d.get('roofline') else None)
print(open('gpurun_out/${tag}_bench_ref.json').read()[:400])
PY
