#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "async or spill or top_degree or multi_source or dblp or hub" > gpurun_out/t9_async.log 2>&1
echo "async tests exit $?" >> gpurun_out/t9_async.log; grep -E "watchdog|passed|failed|Error" gpurun_out/t9_async.log | head -20
for args in "--shape dblp" "--shape youtube" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== probe $args"; timeout 300 python scripts/probe.py $args --show 0 2>&1 | tail -9
done > gpurun_out/t9_probe.log 2>&1
for c in 1 2; do echo "=== youtube ASYNC_CTAS_PER_SM=$c"; DPPR_ASYNC_CTAS_PER_SM=$c timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | tail -6; done >> gpurun_out/t9_probe.log 2>&1
cat gpurun_out/t9_probe.log
