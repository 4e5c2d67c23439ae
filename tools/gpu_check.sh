#!/bin/bash
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
STEPS=10 timeout 300 python tools/role_times.py > gpurun_out/${tag}_roles.txt 2>&1; cat gpurun_out/${tag}_roles.txt
