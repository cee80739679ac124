#!/bin/bash
mkdir -p gpurun_out
for env in "DPPR_RELABEL_BLOCKS=1024" "DPPR_RELABEL_BLOCKS=64" "DPPR_RELABEL_BLOCKS=16384"; do
  for args in "--shape youtube" "--shape orkut --scale 0.25 --batches 20"; do
  echo "=== $env $args"; env $env timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
  done
done
for env in "DPPR_RELABEL_BLOCKS=1024" "DPPR_RELABEL_BLOCKS=16384"; do
env $env timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 5 --kinds rank1k --check 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$env', {k:d[k] for k in ('kind','push_ms_mean','push_edges_per_ns')})"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
