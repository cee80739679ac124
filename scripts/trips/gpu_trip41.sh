#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
echo "=== youtube default"; timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms"
echo "=== youtube forced dense div 8"; DPPR_DENSE_MIN_EDGES=0 DPPR_ITERLOG=1 timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|^\(" | cut -c1-500
for div in 16 64; do
  echo "=== DPPR_DENSE_DIV=$div orkut/4"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-600
done
for div in 16 64; do
  echo "=== DPPR_DENSE_DIV=$div lj/4"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-600
done
