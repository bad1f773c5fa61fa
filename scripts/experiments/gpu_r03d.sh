#!/bin/bash
# the intermittent 2-rank failure of heat1d_small_f_cf2: 12 solves per process, several switches
C="heat1d_small_f_cf2"; L="$C $C $C $C $C $C $C $C $C $C $C $C"
run() { local dir=$1; shift; out=$(cd $dir && env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tests/mp_gpu_case.py $L 2>&1 | grep "^OK\|^FAIL" | cut -c1-4 | sort | uniq -c | tr '\n' ' '); echo "$dir $*: $out"; }
run . A=1
run . MGB_ZERO_U=1
run _old A=1
run . MGB_SINE_MODES=0
run . MGB_LAZY_F=0
run . A=2
