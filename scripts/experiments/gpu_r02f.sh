#!/bin/bash
# 2 GPUs: the 19-case multi-rank parity test, bench at N=2 (parity block), e2e phases at N=1 and N=2
tag=${1:-r02f}
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s > gpurun_out/${tag}_multirank_parity.txt 2>&1; echo "multirank rc=$?"
tail -30 gpurun_out/${tag}_multirank_parity.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['parity'])
PY
tail -3 gpurun_out/${tag}_bench_n2.err
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
head -18 gpurun_out/${tag}_e2e_breakdown_n1.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n2.txt 2>&1
head -18 gpurun_out/${tag}_e2e_breakdown_n2.txt
