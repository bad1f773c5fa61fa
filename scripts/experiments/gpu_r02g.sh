#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
CASES="heat1d_small_v heat1d_small_f_cf2 heat1d_cfg2_nt1025 heat1d_small_jump heat1d_small_tnorminf heat1d_small_weight heat1d_trailing_f dahlquist_cfg1 dahlquist_ml1 advection_example brusselator_example heat1d_example heat1d_bdf2_example heat1d_bdf1_small heat2d_cn_3lvl heat1d_spatial_example heat1d_spatial_large heat1d_atmgrit_k8 heat1d_atmgrit_k5_f"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/mp_gpu_case.py $CASES > gpurun_out/${tag}_multirank_parity.txt 2>&1; echo "rc=$?"
grep "OK \|FAIL" gpurun_out/${tag}_multirank_parity.txt | cut -c1-600
echo "--- node space"
MGB_HEAT1D_SINE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/mp_gpu_case.py $CASES > gpurun_out/${tag}_multirank_parity_node.txt 2>&1; echo "rc=$?"
grep "OK \|FAIL" gpurun_out/${tag}_multirank_parity_node.txt | cut -c1-600
echo "--- no queue ahead / no lazy"
MGB_QUEUE_AHEAD=0 MGB_LAZY_F=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_gpu_case.py heat1d_small_v heat1d_small_f_cf2 heat1d_cfg2_nt1025 heat1d_small_jump heat1d_small_tnorminf heat1d_small_weight heat1d_trailing_f 2>&1 | grep "OK \|FAIL" | cut -c1-600
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
