#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t28_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t28_tests.log; tail -4 gpurun_out/t28_tests.log
for env in "DPPR_COOP_WINDOW=1" "DPPR_COOP_WINDOW=0"; do
  for args in "--shape youtube" "--shape dblp" "--shape orkut --scale 0.25 --batches 20" "--shape livejournal --scale 0.25 --batches 20"; do
  echo "=== $env $args"; env $env timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
  done
done
