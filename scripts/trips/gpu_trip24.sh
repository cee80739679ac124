#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t24_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t24_tests.log; tail -4 gpurun_out/t24_tests.log
for args in "--shape youtube" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100" "--shape orkut --scale 0.25 --batches 20"; do
  echo "=== probe $args"; timeout 300 python scripts/probe.py $args --show 0 2>&1 | tail -5
done > gpurun_out/t24_probe.log 2>&1
grep -E "===|mean ms|per batch|push algo" gpurun_out/t24_probe.log
timeout 1200 python scripts/run_twitter.py --scale 1.0 --batches 10 --top-batches 2 --kinds rank1m,rank1k --check 1 > gpurun_out/t24_tw.jsonl 2> gpurun_out/t24_tw.err
python - <<'PY'
import json
for l in open('gpurun_out/t24_tw.jsonl'):
    d=json.loads(l); print({k:d[k] for k in ('kind','push_ms_mean','step_ms_p50','edge_updates_per_s_step','iterations','traversed','push_edges_per_ns','push_alg_GBps','max_abs_residual_over_eps','invariant_defect','window_checksum_ok')})
PY
