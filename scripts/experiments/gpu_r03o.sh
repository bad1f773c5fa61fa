#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r03_bench_cfg5_n1.json 2> gpurun_out/r03_bench_cfg5_n1.err; echo "bench cfg5 rc=$?"
cut -c1-600 gpurun_out/r03_bench_cfg5_n1.json
