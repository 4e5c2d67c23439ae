#!/bin/bash
mkdir -p gpurun_out
python tools/k1_bench.py 16 64 256 > gpurun_out/k1_bench.txt 2>&1; cat gpurun_out/k1_bench.txt
timeout 600 ncu --set full --clock-control none -k regex:tc_wmsa -s 3 -c 1 -f -o gpurun_out/k1_b16 python tools/k1_bench.py 16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:tc_wmsa -s 3 -c 1 -f -o gpurun_out/k1_b256 python tools/k1_bench.py 256 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_wmsa|tc_resadd2|tc_upconv1p|tc_mlp96" -s 40 -c 8 -f -o gpurun_out/r02_tc python tools/one_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
