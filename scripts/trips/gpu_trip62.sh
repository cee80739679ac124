#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/t62_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t62_tests.log
for args in "--shape youtube --batches 50" "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== $args"; DPPR_ITERLOG=1 DPPR_PROBE_ITER=12 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms|min .* p50" 
done
