#!/bin/bash
mkdir -p gpurun_out
DPPR_CARRY_GAMMA=0.7 timeout 2400 python scripts/run_twitter.py --scale 1.0 --batches 10 --top-batches 3 --kinds rank1m,rank1k > gpurun_out/t23_tw_carry.jsonl 2> gpurun_out/t23_tw_carry.err; tail -2 gpurun_out/t23_tw_carry.err
python - <<'PY'
import json
for l in open('gpurun_out/t23_tw_carry.jsonl'):
    d=json.loads(l); print({k:d[k] for k in ('kind','push_ms_mean','step_ms_p50','edge_updates_per_s_step','iterations','pops','traversed','push_edges_per_ns','max_abs_residual_over_eps','invariant_defect','window_checksum_ok')})
PY
DPPR_CARRY_GAMMA=0.7 timeout 900 python scripts/run_config.py --config 4 --sources 1 --batches 10 --check 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ppr_ms_mean','iterations','pops','traversed','push_edges_per_ns')})"
