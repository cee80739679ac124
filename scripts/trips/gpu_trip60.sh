#!/bin/bash
for div in 1e15 16 6; do
echo "=== youtube dense kernel div $div"; DPPR_ITERLOG=1 DPPR_DENSE_MIN_EDGES=0 DPPR_DENSE_DIV=$div timeout 120 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|per batch:|^\(" | cut -c1-900
done
