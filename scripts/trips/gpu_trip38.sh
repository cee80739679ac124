#!/bin/bash
mkdir -p gpurun_out
for div in 0 8; do
  echo "=== DPPR_DENSE_DIV=$div youtube"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-3000
done
