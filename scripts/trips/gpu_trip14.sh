#!/bin/bash
for h in 64 128 256 512; do echo "=== youtube levelsync hub=$h"; timeout 300 python scripts/probe.py --shape youtube --mode 3 --hub $h --show 0 2>&1 | tail -4; done
for h in 128 512; do echo "=== orkut/4 levelsync hub=$h"; timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --mode 3 --hub $h --show 0 2>&1 | tail -4; done
