#!/bin/bash
# 2 GPUs: the headline workload with the predicted-last-cycle store (parity block), and two small F/V cases
tag=${1:-r03n}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n2.json 2> gpurun_out/${tag}_bench_cfg5_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_cfg5_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['ms_per_step'], d['parity']['ok'], d['parity']['points'], d['parity'].get('conv_abs_diff'), d['config']['conv'])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mp_gpu_case.py heat1d_small_f_cf2 heat1d_cfg2_nt1025 2>&1 | grep "^OK\|^FAIL" | cut -c1-200
