#!/bin/bash
for s in 1 4 16 64; do echo "=== youtube sources=$s"; timeout 600 python scripts/probe.py --shape youtube --sources $s --batches 30 --show 0 2>&1 | tail -5 | head -4; done
