cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pass -s 2 -c 1 -o gpurun_out/t7_eval python tools/eval_bench.py > gpurun_out/t7_ncu.log 2>&1
tail -2 gpurun_out/t7_ncu.log | cut -c1-200
