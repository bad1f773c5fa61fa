#!/bin/bash
# how often does heat1d_small_f_cf2 fail on 2 ranks: current tree vs the tree at the start of the session (_old/)
run() { local dir=$1; shift; local ok=0 fail=0; for i in 1 2 3 4 5 6 7 8; do
  out=$(cd $dir && env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tests/mp_gpu_case.py heat1d_small_f_cf2 2>&1 | grep "^OK\|^FAIL" | cut -c1-60)
  case "$out" in OK*) ok=$((ok+1));; *) fail=$((fail+1));; esac; done; echo "$dir $*: ok=$ok fail=$fail"; }
run . A=1
run _old A=1
run . MGB_SINE_MODES=0
run . MGB_QUEUE_AHEAD=0
run . MGB_PEER_EXCHANGE=0
