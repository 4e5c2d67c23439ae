#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8
for v in on off on off; do
  if [ $v = off ]; then export SJ_WMSA_MAX_C=96; else unset SJ_WMSA_MAX_C; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v  ms_per_step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'])"
done
unset SJ_WMSA_MAX_C
ROLES=enc STEPS=20 timeout 300 python tools/role_times.py 2>&1 | tail -2
