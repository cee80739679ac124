#!/bin/bash
for args in "--shape youtube --batches 50" "--shape livejournal --scale 0.25 --batches 10" "--shape orkut --scale 0.25 --batches 10"; do
  echo "=== switching kernel (3 CTAs/SM) $args"; DPPR_DENSE_MIN_EDGES=0 timeout 120 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
done
timeout 600 python scripts/run_twitter.py --scale 1.0 --batches 3 --top-batches 2 --kinds top,rank1m --check 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('twitter', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags')})"
timeout 600 python scripts/run_twitter.py --V 3072441 --M 117185083 --undirected 1 --batches 2 --top-batches 2 --sources 125 --kinds top --check 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('orkut-size S=125', {k:d.get(k) for k in ('push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags')})"
