#!/bin/bash
# 2 GPUs, light: the F-cycle cases (and a few V-cycle ones) after the dead-F-store change, both row representations
CASES="heat1d_small_f_cf2 heat1d_atmgrit_k5_f heat1d_small_v heat1d_cfg2_nt1025 advection_example heat2d_cn_3lvl heat1d_small_f_cf2 heat1d_small_f_cf2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mp_gpu_case.py $CASES 2>&1 | grep "^OK\|^FAIL" | cut -c1-200
echo "# MGB_HEAT1D_SINE=0"
MGB_HEAT1D_SINE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tests/mp_gpu_case.py heat1d_small_f_cf2 heat1d_atmgrit_k5_f heat1d_small_v 2>&1 | grep "^OK\|^FAIL" | cut -c1-200
