#!/bin/bash
for cap in 1024 512 256 128; do echo "=== orkut/4 TILE_CAP=$cap"; DPPR_TILE_CAP=$cap timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 20 --show 0 2>&1 | grep -E "mean ms|per batch"; done
for cap in 512 256; do echo "=== youtube TILE_CAP=$cap"; DPPR_TILE_CAP=$cap timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | grep -E "mean ms|per batch"; done
for cap in 1024 256; do echo "=== LJ/4 mode1 TILE_CAP=$cap"; DPPR_TILE_CAP=$cap timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --per-batch 100 --batches 100 --show 0 2>&1 | grep -E "mean ms|per batch"; done
for cap in 256; do echo "=== youtube x16 TILE_CAP=$cap"; DPPR_TILE_CAP=$cap timeout 300 python scripts/probe.py --shape youtube --sources 16 --batches 20 --show 0 2>&1 | grep -E "mean ms|per batch"; done
