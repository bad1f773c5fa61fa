#!/bin/bash
# 2 GPUs: the 19 multi-rank parity cases (full log), the pytest wrapper, bench at N=2
tag=${1:-r02}
mkdir -p gpurun_out
CASES="heat1d_small_v heat1d_small_f_cf2 heat1d_cfg2_nt1025 heat1d_small_jump heat1d_small_tnorminf heat1d_small_weight heat1d_trailing_f dahlquist_cfg1 dahlquist_ml1 advection_example brusselator_example heat1d_example heat1d_bdf2_example heat1d_bdf1_small heat2d_cn_3lvl heat1d_spatial_example heat1d_spatial_large heat1d_atmgrit_k8 heat1d_atmgrit_k5_f"
{ echo "# python -m torch.distributed.run --nproc-per-node 2 tests/mp_gpu_case.py <19 cases>  (2 x B200, NCCL + peer-memory mailbox)"; nvidia-smi -L;
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/mp_gpu_case.py $CASES 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$"; echo "rc=$?";
  echo "# the same with the level rows of Heat1D in node space (MGB_HEAT1D_SINE=0)";
  MGB_HEAT1D_SINE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/mp_gpu_case.py $CASES 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$"; echo "rc=$?";
  echo "# python -m pytest tests/test_gpu_multirank.py -m gpu -q";
  timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -3; } > gpurun_out/${tag}_multirank_parity.txt 2>&1
cat gpurun_out/${tag}_multirank_parity.txt | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5_n2.json 2> gpurun_out/${tag}_bench_cfg5_n2.err; echo "bench n2 rc=$?"
cut -c1-900 gpurun_out/${tag}_bench_cfg5_n2.json
