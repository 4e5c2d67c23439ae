cd $GRAFT_REPO_ROOT
MASKS=0,1,2,3,0,3,1,2 timeout 300 python tools/pdl_probe.py 2>&1 | tail -12 | tee gpurun_out/t2_pdl_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_gemm_kernel<2, false' -c 4 -o gpurun_out/t2_resadd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t2_ncu.log 2>&1
tail -5 gpurun_out/t2_ncu.log
