set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/t1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1_pytest.log
tail -25 gpurun_out/t1_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/t1_bench_pdl.json 2> gpurun_out/t1_bench_pdl.err; echo rc=$?
SJ_NO_PDL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t1_bench_nopdl.json 2> gpurun_out/t1_bench_nopdl.err; echo rc=$?
SJ_NO_RPF=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t1_bench_norpf.json 2> gpurun_out/t1_bench_norpf.err; echo rc=$?
cat gpurun_out/t1_bench_pdl.json | cut -c1-400; cat gpurun_out/t1_bench_nopdl.json | cut -c1-300; cat gpurun_out/t1_bench_norpf.json | cut -c1-300
