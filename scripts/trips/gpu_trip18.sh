#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/t18_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t18_tests.log; tail -4 gpurun_out/t18_tests.log
for env in "DPPR_FUSED_WINDOW=1" "DPPR_FUSED_WINDOW=0"; do
  for args in "--shape youtube" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== $env $args"; env $env timeout 300 python scripts/probe.py $args --show 0 2>&1 | tail -6 | head -3
  done
done > gpurun_out/t18_probe.log 2>&1
cat gpurun_out/t18_probe.log
