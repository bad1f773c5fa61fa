#!/bin/bash
# 8 GPUs: bench at N = 8, 4, 2 (parity block in every line), timeline and e2e phases at N = 8
tag=${1:-r02y}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err; echo "bench n$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench_n${n}.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['parity']['ok'], d['parity']['points'], d['parity'].get('conv_abs_diff'), d['config']['conv'])
except Exception as e: print('failed', e)
PY
  tail -2 gpurun_out/${tag}_bench_n${n}.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline_n8.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${tag}_timeline_n8.txt | head -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n8.txt 2>&1
grep -A1 "^hierarchy" gpurun_out/${tag}_e2e_breakdown_n8.txt | head -18
timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['parity']['ok'])
PY
MGRIT_BENCH_SMI_MS=200 timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/${tag}_bench_n1_smi200.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_n1_smi200.json'))
print('smi every 200 ms:', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['samples_ms'])
PY
