#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "n$N rc=$?"; cat gpurun_out/n${N}_bench.json | cut -c1-400; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n${N}_bench.json"))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"fp32io",d["e2e"]["fp32_io"]["value"],"dp",d.get("dp_invariant_ok"),d["config"]["collective"][:60])
except Exception as e:
    print("parse failed",e)
PY
tail -5 gpurun_out/n${N}_bench.err
