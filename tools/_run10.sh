cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo ref rc=$?
timeout 600 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo bench rc=$?
cut -c1-260 gpurun_out/f_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python tools/one_step.py > gpurun_out/f_one_step.log 2>&1; tail -1 gpurun_out/f_one_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_wmsa|tc_mlp96|tc_upconv|tc_outconv' -s 30 -c 15 -o gpurun_out/f_tc python tools/one_step.py > gpurun_out/f_ncu_tc.log 2>&1; tail -2 gpurun_out/f_ncu_tc.log
python __graft_entry__.py smoke 2>&1 | tail -4
ls -la gpurun_out
