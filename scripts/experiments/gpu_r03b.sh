#!/bin/bash
# bisect heat1d_small_f_cf2 on 2 ranks
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tests/mp_gpu_case.py heat1d_small_f_cf2 heat1d_small_v 2>&1 | grep "^OK\|^FAIL" | cut -c1-200; }
run A=1
run MGB_SINE_MODES=0
run MGB_QUEUE_AHEAD=0
run MGB_LAZY_F=0
run MGB_FUSED_DOWN=0
run MGB_PEER_EXCHANGE=0
run A=2
