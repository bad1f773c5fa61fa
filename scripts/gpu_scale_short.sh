#!/bin/bash
# 8 GPUs, short: bench at N = 8 and 4 (parity block in every line), timeline and e2e phases at N = 8
tag=${1:-r03}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n${n}.json 2> gpurun_out/${tag}_bench_cfg5_n${n}.err; echo "bench n$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench_cfg5_n${n}.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['ms_per_step'], d['e2e']['samples_ms'], d['parity']['ok'], d['parity']['points'], d['parity'].get('conv_abs_diff'), d['config']['conv'])
except Exception as e: print('failed', e)
PY
  tail -2 gpurun_out/${tag}_bench_cfg5_n${n}.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline_cfg5_n8.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${tag}_timeline_cfg5_n8.txt | head -24
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n8.txt 2>&1
grep -A1 "^hierarchy" gpurun_out/${tag}_e2e_breakdown_n8.txt | head -12
