#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
for both in 1 0; do
DPPR_RELABEL_BOTH=$both timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 4 --top-batches 2 --kinds top,rank1k,rank1m --check 0 2>gpurun_out/t59_tw_$both.err | tee gpurun_out/t59_tw_$both.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('both=$both', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','dense_sweeps','push_edges_per_ns','push_ms_each','error_flags')})"
done
