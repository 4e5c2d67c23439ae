cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/t4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t4_pytest.log
tail -30 gpurun_out/t4_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err; echo rc=$?
cut -c1-330 gpurun_out/t4_bench.json
