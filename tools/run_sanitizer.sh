#!/bin/bash
mkdir -p gpurun_out
timeout 2000 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/sanitizer_memcheck_all.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitizer_memcheck_all.log | head -20
