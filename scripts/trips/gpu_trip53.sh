#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -4
for args in "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10"; do
  echo "=== dense kernel $args"; DPPR_ITERLOG=1 DPPR_DENSE_MIN_EDGES=0 timeout 120 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms|Error|error|^\(" | cut -c1-600
done
for h in 3 0 1 2; do
lib=$PWD/dynamicppr_b200/lib/libdppr_h$h.so; [ $h = 3 ] && lib=$PWD/dynamicppr_b200/lib/libdppr.so
DPPR_LIB=$lib timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 3 --top-batches 3 --kinds top --check 0 2>gpurun_out/t53_tw_$h.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('hints=$h', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','dense_sweeps','push_ms_each','error_flags')})"
done
