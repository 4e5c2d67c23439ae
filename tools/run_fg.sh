#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider -k "fgmsa or strajnet or smoke or forward" 2>&1 | tail -4
ROLES=fgmsa STEPS=20 timeout 300 python tools/role_times.py 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  ms_per_step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'])"
done
