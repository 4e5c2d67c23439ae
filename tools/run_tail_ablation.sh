#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/tail_ablation3.txt
: > $out
for d in 170 186 426 682 954; do
  SJ_UP4H_DBG=$d timeout 120 python tools/tail_bench.py 16 >> $out 2>&1 || echo "dbg=$d failed" >> $out
done
cat $out
