#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
for div in 16; do
  echo "=== DPPR_DENSE_DIV=$div orkut/4"; DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape orkut --scale 0.25 --batches 10 --show 0 2>&1 | grep -E "mean ms"
done
for div in 8 32; do
DPPR_DENSE_DIV=$div timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 4 --kinds rank1k,rank1m --check 1 2>gpurun_out/t42_tw_$div.err | tee gpurun_out/t42_tw_$div.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('div=$div', {k:d.get(k) for k in ('kind','push_ms_mean','window_ms_mean','step_ms_p50','iterations','traversed','push_edges_per_ns','max_abs_residual_over_eps','invariant_defect','window_checksum_ok','csr_entries_ok','error_flags')})"
done
