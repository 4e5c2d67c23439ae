#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in 0 1; do
  if [ $v = 1 ]; then export SJ_SIDE_STREAM=1; else unset SJ_SIDE_STREAM; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/side_$v.json 2> gpurun_out/side_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/side_$v.json"))
print("rep $rep SJ_SIDE_STREAM=$v ms/step graph %.4f stream %.4f e2e %.4f"%(d["ms_per_step"], d["ms_per_step_stream_launches"], d["e2e"]["ms_per_step"]))
PY
done
done
