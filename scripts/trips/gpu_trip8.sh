#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "async or spill or top_degree or multi_source or dblp" > gpurun_out/t8_async.log 2>&1
echo "async tests exit $?" >> gpurun_out/t8_async.log; tail -15 gpurun_out/t8_async.log
for args in "--shape dblp" "--shape youtube" "--shape youtube --mode 3" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100" "--shape orkut --scale 0.25 --batches 20"; do
  echo "=== probe $args"; timeout 300 python scripts/probe.py $args --show 0 2>&1 | tail -9
done > gpurun_out/t8_probe.log 2>&1
cat gpurun_out/t8_probe.log
