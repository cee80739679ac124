#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/t47_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/t47_tests.log
for args in "--shape youtube --batches 50" "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10"; do
  echo "=== $args"; timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
done
for div in 16 32 64; do
echo "=== orkut/4 div $div"; DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
echo "=== lj/4 div $div"; DPPR_DENSE_MIN_EDGES=0 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
done
for div in 32 128; do
DPPR_DENSE_DIV=$div timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 3 --top-batches 2 --kinds rank1k,top,rank1m --check 0 2>gpurun_out/t47_tw_$div.err | tee gpurun_out/t47_tw_$div.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('div=$div', {k:d.get(k) for k in ('kind','push_ms_mean','step_ms_p50','iterations','push_edges_per_ns','error_flags')})"
done
