#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the bench command, ncu full capture of level-0 sweeps.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
# second pass of scripts/profile_sweeps.py: f_relax, f_relax(last), c_relax, fas_residual, down_sweep, correct+F, residual
COARSENING=64 LEVELS=3 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_c_relax|k_fas_residual|k_down|k_correct|k_residual" -s 7 -c 7 \
    -o gpurun_out/${tag}_sweeps -f python scripts/profile_sweeps.py > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/solve_timeline.py cfg5 > gpurun_out/${tag}_timeline.txt 2>&1
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown.txt 2>&1
ls -la gpurun_out
