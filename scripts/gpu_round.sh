#!/bin/bash
# One GPU-box visit (1 GPU): parity tests, bench lines of every BASELINE config, reference arm, ncu launch list of the bench
# command, ncu full capture of the level-0 sweeps, timeline, e2e phases.
# Usage (from the repo root): gpurun --timeout 3000 -- 'bash scripts/gpu_round.sh <tag>'
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n1.json 2> gpurun_out/${tag}_bench_cfg5_n1.err; echo "bench cfg5 rc=$?"
cut -c1-1200 gpurun_out/${tag}_bench_cfg5_n1.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench_cfg5_n1.err
cut -c1-600 gpurun_out/${tag}_bench_ref.json
for wl in cfg1 cfg2 cfg3; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${wl}_n1.json 2> gpurun_out/${tag}_bench_${wl}_n1.err; echo "bench $wl rc=$?"
done
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg4_n1.json 2> gpurun_out/${tag}_bench_cfg4_n1.err; echo "bench cfg4 rc=$?"
python - <<PY
import json
for wl in ('cfg5','cfg1','cfg2','cfg3','cfg4'):
    try:
        d=json.load(open('gpurun_out/${tag}_bench_%s_n1.json' % wl))
        print(wl, 'device %.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], 'cold %.1f' % d['e2e']['cold_ms'], 'iters', d['config']['iterations'], 'launches', d['gpu_launches'], 'parity', d['parity'] and d['parity']['ok'], 'cpu', d['cpu_baseline'] and (d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind']))
        for k in d['kernels']: print('   %-45s %.3f ms  hbm %.2f  fp64 %s  x%d/it x%d/solve' % (k['name'],k['ms'],k.get('hbm_frac',0),('%.2f' % k['fp64_frac']) if 'fp64_frac' in k else '-',k['launches_per_iteration'],k['launches_per_solve']))
    except Exception as e: print(wl, 'failed', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual" -s 5 -c 5 \
   -o gpurun_out/${tag}_sweeps -f python scripts/profile_sine.py > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline_cfg5_n1.txt 2>&1; head -18 gpurun_out/${tag}_timeline_cfg5_n1.txt
MGB_HEAT1D_SINE=0 python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline_cfg5_n1_node_space.txt 2>&1; head -3 gpurun_out/${tag}_timeline_cfg5_n1_node_space.txt
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1; head -8 gpurun_out/${tag}_e2e_breakdown_n1.txt
for wl in cfg2 cfg3 cfg4; do python scripts/solve_timeline.py $wl > gpurun_out/${tag}_timeline_${wl}_n1.txt 2>&1; head -3 gpurun_out/${tag}_timeline_${wl}_n1.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_rows_rfft|k_rows_irfft|k_cplx_solve" -s 3 -c 3 \
   -o gpurun_out/${tag}_fourier -f python scripts/profile_fourier.py > gpurun_out/${tag}_ncu_fourier.log 2>&1; echo "ncu fourier rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual" -s 5 -c 5 \
   -o gpurun_out/${tag}_cfg3_modes -f python scripts/profile_cfg3.py > gpurun_out/${tag}_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
ls gpurun_out | grep ${tag}
