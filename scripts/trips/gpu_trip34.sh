#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t34_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t34_tests.log; tail -4 gpurun_out/t34_tests.log
for env in "DPPR_RELABEL=1" "DPPR_RELABEL=0"; do
  for args in "--shape youtube" "--shape orkut --scale 0.25 --batches 20" "--shape livejournal --scale 0.25 --per-batch 100 --batches 100"; do
  echo "=== $env $args"; env $env timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms|per batch"
  done
done
timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 10 --top-batches 2 --kinds rank1m,rank1k > gpurun_out/t34_tw.jsonl 2> gpurun_out/t34_tw.err
python - <<'PY'
import json
for l in open('gpurun_out/t34_tw.jsonl'):
    d=json.loads(l); print({k:d[k] for k in ('kind','push_ms_mean','step_ms_p50','edge_updates_per_s_step','iterations','traversed','push_edges_per_ns','max_abs_residual_over_eps','invariant_defect','window_checksum_ok','csr_entries_ok','window_ms_mean','init_window_s')})
PY
