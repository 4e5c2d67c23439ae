cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_eval.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/eval_bench.py 2>&1 | tail -3 | tee gpurun_out/t6_eval_bench.json
