#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -4
echo "=== youtube default"; timeout 300 python scripts/probe.py --shape youtube --batches 50 --show 0 2>&1 | grep -E "mean ms"
echo "=== youtube dense kernel"; DPPR_DENSE_MIN_EDGES=0 timeout 300 python scripts/probe.py --shape youtube --batches 50 --show 0 2>&1 | grep -E "mean ms"
for args in "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10"; do
  echo "=== dense kernel $args"; DPPR_DENSE_MIN_EDGES=0 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
  echo "=== off $args"; DPPR_DENSE_DIV=0 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
done
DPPR_ITERLOG=1 timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 4 --top-batches 2 --kinds rank1k,top,rank1m --check 0 2>gpurun_out/t52_tw.err | tee gpurun_out/t52_tw.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('kind','push_ms_mean','step_ms_p50','iterations','dense_sweeps','push_ms_each','push_edges_per_ns','error_flags')})"
grep "per-iteration" gpurun_out/t52_tw.err | cut -c1-500
