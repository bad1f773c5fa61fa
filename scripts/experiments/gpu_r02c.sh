#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_chain|k_down|k_correct|k_residual" -s 5 -c 5 \
   -o gpurun_out/${tag}_sine -f python scripts/profile_sine.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}*
