#!/bin/bash
echo "=== youtube default"; timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | grep -E "mean ms|per batch"
for cap in 128 64 256; do echo "=== orkut/4 TILE_CAP=$cap"; DPPR_TILE_CAP=$cap timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 20 --show 0 2>&1 | grep -E "mean ms|per batch"; done
echo "=== LJ/4 mode1"; timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --per-batch 100 --batches 100 --show 0 2>&1 | grep -E "mean ms|per batch"
echo "=== youtube x16"; timeout 300 python scripts/probe.py --shape youtube --sources 16 --batches 20 --show 0 2>&1 | grep -E "mean ms|per batch"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
