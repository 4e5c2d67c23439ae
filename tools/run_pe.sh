#!/bin/bash
# fused patch embedding vs the im2col -> GEMM -> combine chain: role times without the actor side stream, then the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k patch_embed -p no:cacheprovider 2>&1 | tail -5
for v in fused chain; do
  if [ $v = chain ]; then export SJ_DISABLE_FUSED_PE=1; else unset SJ_DISABLE_FUSED_PE; fi
  echo "== $v"
  SJ_NO_SIDE_STREAM=1 ROLES=enc.pe,enc STEPS=20 timeout 300 python tools/role_times.py 2>&1 | tail -3
  for i in 1 2; do
    timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  ms_per_step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'])"
  done
done
unset SJ_DISABLE_FUSED_PE
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_patch_embed_kernel -s 4 -c 2 -f -o gpurun_out/pe_fused python tools/one_step.py > gpurun_out/pe_fused.log 2>&1
ls -la gpurun_out/pe_fused.ncu-rep
