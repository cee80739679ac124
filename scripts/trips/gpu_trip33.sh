#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 10 --top-batches 3 --kinds rank1m,rank1k,top > gpurun_out/t33_tw.jsonl 2> gpurun_out/t33_tw.err
python - <<'PY'
import json
for l in open('gpurun_out/t33_tw.jsonl'):
    d=json.loads(l); print({k:d[k] for k in ('kind','push_ms_mean','step_ms_p50','edge_updates_per_s_step','iterations','traversed','push_edges_per_ns','push_alg_GBps','max_abs_residual_over_eps','invariant_defect','window_checksum_ok','window_ms_mean','repair_ms_mean')})
PY
timeout 1500 python scripts/run_config.py --config 3 --variant 0 --check 1 2>/dev/null | tee gpurun_out/t33_c3.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('config','ppr_ms_p50','edge_updates_per_s','traversed','push_edges_per_ns','window_bit_exact','max_abs_residual_over_eps','invariant_defect')})"
