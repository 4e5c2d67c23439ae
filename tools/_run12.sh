cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/f_bench_n8.json 2> gpurun_out/f_bench_n8.err; echo rc=$?
cut -c1-260 gpurun_out/f_bench_n8.json; tail -3 gpurun_out/f_bench_n8.err
