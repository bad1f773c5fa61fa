#!/bin/bash
tag=${1:-r02u}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_rows_rfft|k_rows_irfft" -s 2 -c 2 \
   -o gpurun_out/${tag}_fourier -f python scripts/profile_fourier.py > gpurun_out/${tag}_ncu_fourier.log 2>&1
echo "ncu fourier rc=$?"; tail -2 gpurun_out/${tag}_ncu_fourier.log
