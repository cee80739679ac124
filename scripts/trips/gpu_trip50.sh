#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -4
for div in 0 3 6; do
echo "=== youtube div $div"; DPPR_ITERLOG=1 DPPR_DENSE_MIN_EDGES=0 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|^\(" | cut -c1-700
echo "=== orkut/4 div $div"; DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|^\(" | cut -c1-500
echo "=== lj/4 div $div"; DPPR_ITERLOG=1 DPPR_DENSE_MIN_EDGES=0 DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape livejournal --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms|^\(" | cut -c1-500
done
for div in 3 6; do
DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 4 --top-batches 2 --kinds rank1k,top,rank1m --check 0 2>gpurun_out/t50_tw_$div.err | tee gpurun_out/t50_tw_$div.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('div=$div', {k:d.get(k) for k in ('kind','push_ms_mean','step_ms_p50','iterations','dense_sweeps','push_edges_per_ns','error_flags')})"
grep "per-iteration" gpurun_out/t50_tw_$div.err | cut -c1-600
done
