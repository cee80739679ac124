#!/bin/bash
for cps in 2 3 4 5 6 8; do
echo "=== youtube ctas/sm=$cps"; DPPR_CTAS_PER_SM=$cps timeout 120 python scripts/probe.py --shape youtube --batches 50 --show 0 2>&1 | grep -E "mean ms|per batch"
done
for hub in 32 128; do
echo "=== youtube hub=$hub"; DPPR_HUB_DEGREE=$hub timeout 120 python scripts/probe.py --shape youtube --batches 50 --show 0 2>&1 | grep -E "mean ms"
done
