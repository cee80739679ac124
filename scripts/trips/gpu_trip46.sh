#!/bin/bash
mkdir -p gpurun_out
for v in B C; do
  L2=$PWD/dynamicppr_b200/lib/libdppr_ipt2$v.so
  for args in "--shape youtube --batches 50" "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== [$v] $args"; DPPR_LIB=$L2 DPPR_DENSE_DIV=0 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
  done
done
