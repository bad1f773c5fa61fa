#!/bin/bash
# Full GPU suite, cfg4/cfg3 bench lines, ncu of the Fourier kernels and of the Heat2D mode kernels, e2e marks.
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --workload cfg4 > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err; echo "bench cfg4 rc=$?"
cut -c1-1200 gpurun_out/${tag}_bench_cfg4.json
timeout 300 python scripts/e2e_breakdown.py cfg5 > gpurun_out/${tag}_e2e_breakdown_n1.txt 2>&1
sed -n '/marks of one run/,/function calls/p' gpurun_out/${tag}_e2e_breakdown_n1.txt | head -60
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_rows_rfft|k_rows_irfft|k_cplx_solve" -s 3 -c 3 \
   -o gpurun_out/${tag}_fourier -f python scripts/profile_fourier.py > gpurun_out/${tag}_ncu_fourier.log 2>&1
echo "ncu fourier rc=$?"; tail -2 gpurun_out/${tag}_ncu_fourier.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual" -s 5 -c 5 \
   -o gpurun_out/${tag}_cfg3_modes -f python scripts/profile_cfg3.py > gpurun_out/${tag}_ncu_cfg3.log 2>&1
echo "ncu cfg3 rc=$?"; tail -2 gpurun_out/${tag}_ncu_cfg3.log
ls -la gpurun_out/*.ncu-rep
