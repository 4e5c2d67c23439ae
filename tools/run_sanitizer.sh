#!/bin/bash
# compute-sanitizer memcheck over the tight per-kernel tests of the kernels added last + one whole forward (smoke)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "decoder_tail or res_add2 or out_heads or patch_embed or (tc_wmsa and 192)" > gpurun_out/sanitizer_kernels.log 2>&1
echo "memcheck tight kernel tests rc=$?"; tail -6 gpurun_out/sanitizer_kernels.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -6 gpurun_out/sanitizer_smoke.log
