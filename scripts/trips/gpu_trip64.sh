#!/bin/bash
mkdir -p gpurun_out
DPPR_ITERLOG=1 timeout 900 python scripts/run_twitter.py --V 3072441 --M 117185083 --undirected 1 --batches 2 --top-batches 2 --sources 125 --kinds top --check 0 2>gpurun_out/t64.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('kind','push_ms_mean','repair_ms_mean','window_ms_mean','iterations','dense_sweeps','push_edges_per_ns','push_ms_each','error_flags')})"
grep "per-iteration" gpurun_out/t64.err | cut -c1-6000
