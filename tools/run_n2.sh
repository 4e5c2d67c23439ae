#!/bin/bash
mkdir -p gpurun_out
timeout 200 tools/bin/mma_probe power > gpurun_out/mma_probe_power.txt 2>&1; cat gpurun_out/mma_probe_power.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "n2 rc=$?"; cat gpurun_out/n2_bench.json; tail -5 gpurun_out/n2_bench.err
