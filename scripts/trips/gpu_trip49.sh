#!/bin/bash
mkdir -p gpurun_out
for div in 0 3; do
DPPR_ITERLOG=1 DPPR_DENSE_DIV=$div timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 6 --kinds rank1m --check 0 2>gpurun_out/t49_tw_$div.err | tee gpurun_out/t49_tw_$div.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('div=$div', {k:d.get(k) for k in ('kind','push_ms_mean','step_ms_p50','iterations','dense_sweeps','push_ms_each','push_edges_per_ns','traversed','error_flags')})"
grep "per-iteration" gpurun_out/t49_tw_$div.err | cut -c1-2500
done
