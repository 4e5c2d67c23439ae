#!/bin/bash
# fused patch embedding: role times without the actor side stream for a few grid sizes, then the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k patch_embed -p no:cacheprovider 2>&1 | tail -5
for g in 0 296 148 444; do
  echo "== SJ_PE_GRID=$g"
  SJ_PE_GRID=$g SJ_NO_SIDE_STREAM=1 ROLES=enc.pe STEPS=20 timeout 300 python tools/role_times.py 2>&1 | tail -2 | head -1
done
for i in 1 2; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  ms_per_step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'])"
done
