#!/bin/bash
mkdir -p gpurun_out
N=$(python tools/one_step.py 2>/dev/null | grep -o "[0-9]*$")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_gemm_kernel -c 400 --csv --log-file gpurun_out/gemm_list.csv python tools/one_step.py > /dev/null 2>&1
G=$(grep -c tc_gemm_kernel gpurun_out/gemm_list.csv); PER=$((G/3)); echo "tc_gemm launches per forward: $PER"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s $((2*PER+5)) -c 8 -f -o gpurun_out/gemm8 python tools/one_step.py > gpurun_out/gemm8.log 2>&1
ls -la gpurun_out/gemm8.ncu-rep
