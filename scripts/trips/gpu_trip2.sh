#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 > gpurun_out/t2_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/t2_parity.log
tail -15 gpurun_out/t2_parity.log
for args in "--shape dblp --source-kind low" "--shape dblp" "--shape youtube" "--shape youtube --mode 1" "--shape youtube --variant 1" "--shape youtube --variant 2" "--shape youtube --variant 3" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== probe $args"; timeout 600 python scripts/probe.py $args --show 2 2>&1 | tail -14
done > gpurun_out/t2_probe.log 2>&1
cat gpurun_out/t2_probe.log
