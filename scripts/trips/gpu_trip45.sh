#!/bin/bash
mkdir -p gpurun_out
L2=$PWD/dynamicppr_b200/lib/libdppr_ipt2.so
for lib in "" "DPPR_LIB=$L2"; do
  for args in "--shape youtube --batches 50" "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== [$lib] $args"; env $lib DPPR_DENSE_DIV=0 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
  done
done
echo "=== ipt2 8 CTAs/SM orkut/4"; DPPR_LIB=$L2 DPPR_CTAS_PER_SM=8 DPPR_DENSE_DIV=0 timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
echo "=== ipt2 6 CTAs/SM orkut/4"; DPPR_LIB=$L2 DPPR_CTAS_PER_SM=6 DPPR_DENSE_DIV=0 timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
echo "=== ipt2 dense div 16 orkut/4"; DPPR_LIB=$L2 DPPR_DENSE_DIV=16 timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
