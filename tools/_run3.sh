cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_gemm_kernel<\(int\)2, \(bool\)0' -c 3 -o gpurun_out/t3_resadd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t3_ncu.log 2>&1
tail -3 gpurun_out/t3_ncu.log | cut -c1-300
ls -la gpurun_out/
