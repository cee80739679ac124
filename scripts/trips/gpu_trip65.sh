#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
DPPR_ITERLOG=1 timeout 900 python scripts/run_twitter.py --V 3072441 --M 117185083 --undirected 1 --batches 3 --top-batches 3 --sources 125 --kinds top --check 0 2>gpurun_out/t65.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('orkut-size S=125', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','dense_sweeps','push_edges_per_ns','push_ms_each','error_flags')})"
grep "per-iteration" gpurun_out/t65.err | cut -c1-400
timeout 900 python scripts/run_twitter.py --scale 1.0 --batches 3 --top-batches 3 --kinds top,rank1k --check 1 2>gpurun_out/t65b.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('twitter', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags','max_abs_residual_over_eps','invariant_defect','window_checksum_ok')})"
