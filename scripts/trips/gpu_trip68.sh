#!/bin/bash
mkdir -p gpurun_out
for cps in 4 6 8; do
DPPR_CTAS_PER_SM=$cps timeout 600 python scripts/run_twitter.py --V 3072441 --M 117185083 --undirected 1 --batches 2 --top-batches 2 --sources 125 --kinds top --check 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('orkut-size S=125 ctas/sm=$cps', {k:d.get(k) for k in ('push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags')})"
done
for cps in 6 8; do
DPPR_CTAS_PER_SM=$cps timeout 600 python scripts/run_twitter.py --scale 1.0 --batches 2 --top-batches 2 --kinds top --check 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('twitter top ctas/sm=$cps', {k:d.get(k) for k in ('push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags')})"
done
