#!/bin/bash
C="heat1d_small_f_cf2"; L="$C $C $C $C $C $C $C $C $C $C $C $C $C $C $C $C"
run() { out=$(env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tests/mp_gpu_case.py $L 2>&1 | grep "^OK\|^FAIL" | cut -c1-4 | sort | uniq -c | tr '\n' ' '); echo "$*: $out"; }
run MGB_SINE_MODES_MASK=14
run MGB_SINE_MODES_MASK=13
run MGB_SINE_MODES_MASK=11
run MGB_SINE_MODES_MASK=7
run MGB_SINE_MODES_MASK=15
