#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x > gpurun_out/t4_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/t4_parity.log
tail -8 gpurun_out/t4_parity.log
for args in "--shape dblp" "--shape youtube" "--shape youtube --mode 1" "--shape youtube --variant 1" "--shape youtube --variant 2" "--shape youtube --variant 3" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== probe $args"; timeout 600 python scripts/probe.py $args --show 0 2>&1 | tail -9
done > gpurun_out/t4_probe.log 2>&1
for c in 1 2 3; do echo "=== youtube CTAS_PER_SM=$c"; DPPR_CTAS_PER_SM=$c timeout 600 python scripts/probe.py --shape youtube --show 0 2>&1 | tail -6; done >> gpurun_out/t4_probe.log 2>&1
cat gpurun_out/t4_probe.log
