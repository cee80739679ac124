#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/t55_c4.jsonl
for S in 1 8; do for div in 4 0; do
  chk=0; [ $div = 4 ] && chk=1
  DPPR_DENSE_DIV=$div timeout 900 python scripts/run_config.py --config 4 --sources $S --batches 5 --check $chk 2>gpurun_out/t55_c4_${S}_$div.err | tee -a gpurun_out/t55_c4.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('S=$S div=$div', {k:d.get(k) for k in ('ppr_ms_mean','window_ms_mean','iterations','dense_sweeps','push_edges_per_ns','source_edge_updates_per_s','error_flags','window_bit_exact','max_abs_residual_over_eps','invariant_defect')})"
done; done
