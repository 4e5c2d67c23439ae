#!/bin/bash
mkdir -p gpurun_out
for m in 0 7 5 6; do
  SJ_PDL_MASK=$m timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/pdl_$m.json 2> gpurun_out/pdl_$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/pdl_$m.json"))
print("SJ_PDL_MASK=$m ms/step graph %.4f stream %.4f e2e %.4f"%(d["ms_per_step"], d["ms_per_step_stream_launches"], d["e2e"]["ms_per_step"]))
PY
done
