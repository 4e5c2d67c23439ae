#!/bin/bash
# final measurement set of the round: GPU tests, smoke, every single-GPU bench configuration, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/final_smoke.log
for cfg in 3 5 2 1; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/final_bench_c${cfg}.json 2> gpurun_out/final_bench_c${cfg}.err
  echo "bench config $cfg rc=$?"; cut -c1-260 gpurun_out/final_bench_c${cfg}.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/final_bench_ref.json
STEPS=10 timeout 300 python tools/role_times.py > gpurun_out/final_roles.txt 2>&1
# launch list of one forward (the third: warm) and full captures of the kernels that are new since the last committed set
N=$(python tools/one_step.py 2>/dev/null | grep -o "[0-9]*$")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python tools/one_step.py > /dev/null 2>&1
python tools/last_step.py gpurun_out/final_launches.csv $N > gpurun_out/final_launches_one_step.txt; tail -25 gpurun_out/final_launches_one_step.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_patch_embed|tc_wmsa|fg_offset_mma|head_tapsum" -s 16 -c 8 -f -o gpurun_out/final_new_kernels python tools/one_step.py > gpurun_out/final_ncu.log 2>&1
ls -la gpurun_out/final_new_kernels.ncu-rep
