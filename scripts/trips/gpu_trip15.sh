#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/t15_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t15_tests.log; tail -4 gpurun_out/t15_tests.log
for env in "DPPR_CARRY_GAMMA=1.0" "DPPR_CARRY_GAMMA=0.7" "DPPR_CARRY_GAMMA=0.8" "DPPR_CARRY_GAMMA=0.6" "DPPR_CARRY_GAMMA=0.7 DPPR_HUB_DEGREE=32" "DPPR_CARRY_GAMMA=0.7 DPPR_HUB_DEGREE=128" "DPPR_CARRY_GAMMA=0.7 DPPR_CARRY_SCALE=0.1" "DPPR_CARRY_GAMMA=0.7 DPPR_CTAS_PER_SM=2" ; do
  echo "=== youtube $env"; env $env timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | tail -5
done > gpurun_out/t15_probe.log 2>&1
for args in "--shape dblp" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100" "--shape orkut --scale 0.25 --batches 20"; do
  echo "=== probe $args"; timeout 300 python scripts/probe.py $args --show 0 2>&1 | tail -5
done >> gpurun_out/t15_probe.log 2>&1
grep -E "===|mean ms|per batch|push algo" gpurun_out/t15_probe.log
