#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -6
for S in 8 32; do for grp in 1 8; do
  echo "=== orkut/4 sources $S group $grp"; DPPR_PULL_GROUP=$grp timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 5 --sources $S --show 0 2>&1 | grep -E "mean ms|per batch"
done; done
