cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_n2_ce.json 2> gpurun_out/g_bench_n2_ce.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/g_bench_n2_ce.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['collective'])"
grep -i "warn\|error\|Traceback" -A3 gpurun_out/g_bench_n2_ce.err | head -20
SJ_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_n2_nccl.json 2> gpurun_out/g_bench_n2_nccl.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/g_bench_n2_nccl.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['collective'])"
