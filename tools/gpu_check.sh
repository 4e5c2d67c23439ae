#!/bin/bash
# usage: tools/gpu_check.sh TAG [pytest-args...]   -- GPU tests + a short bench, logs under gpurun_out/TAG_*
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
