#!/bin/bash
# GPU tests (selected files) + bench lines for every single-GPU configuration
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -q -m gpu -p no:cacheprovider -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|rc=|max \|err\||bf16 config|Error|error" gpurun_out/${tag}_pytest.log | tail -30
for cfg in 3 5 2 1; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c${cfg}.json 2> gpurun_out/${tag}_bench_c${cfg}.err
  echo "bench config $cfg rc=$?"; cut -c1-600 gpurun_out/${tag}_bench_c${cfg}.json; tail -3 gpurun_out/${tag}_bench_c${cfg}.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${tag}_bench_ref.json
