#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
for div in 0 0.001 8; do
  echo "=== DPPR_DENSE_DIV=$div youtube"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-1200
done
for div in 0 8; do
  echo "=== DPPR_DENSE_DIV=$div orkut/4"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-1500
done
