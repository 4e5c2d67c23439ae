#!/bin/bash
# runs every section of the tcgen05 probe in its own process (a trap in one section must not hide the others)
mkdir -p gpurun_out
out=gpurun_out/mma_probe.txt
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
for sec in tscheck n pattern; do
  echo "== section $sec" >> $out
  timeout 120 tools/bin/mma_probe $sec >> $out 2>&1 || echo "section $sec failed rc=$?" >> $out
done
cat $out
