#!/bin/bash
mkdir -p gpurun_out
for div in 16 64; do
  echo "=== DPPR_DENSE_DIV=$div orkut/4"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-900
done
for div in 0 16 64; do
  echo "=== DPPR_DENSE_DIV=$div lj/4"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-900
done
