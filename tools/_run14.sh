cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_n8_ce.json 2> gpurun_out/g_bench_n8_ce.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/g_bench_n8_ce.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['collective'])"
grep -i "warn\|error\|Traceback" -A3 gpurun_out/g_bench_n8_ce.err | head -20
