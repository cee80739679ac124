#!/bin/bash
for st in 1 2 3 4; do echo "=== youtube SUBTILES=$st"; DPPR_SUBTILES=$st timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | grep -E "mean ms|per batch"; done
for st in 1 2 4; do echo "=== orkut/4 SUBTILES=$st"; DPPR_SUBTILES=$st timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 20 --show 0 2>&1 | grep -E "mean ms|per batch"; done
