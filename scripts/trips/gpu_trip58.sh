#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 3 --top-batches 2 --sources 8 --kinds rank1k,top --check 0 2>gpurun_out/t58_tw8.err | tee gpurun_out/t58_tw8.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('twitter S=8', {k:d.get(k) for k in ('kind','push_ms_mean','window_ms_mean','repair_ms_mean','iterations','dense_sweeps','push_edges_per_ns','edge_updates_per_s_step','error_flags')})"
tail -2 gpurun_out/t58_tw8.err | cut -c1-300
timeout 1500 python scripts/run_config.py --config 4 --sources 125 --batches 3 --check 0 2>gpurun_out/t58_c4_125.err | tee gpurun_out/t58_c4_125.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('orkut S=125', {k:d.get(k) for k in ('ppr_ms_mean','window_ms_mean','repair_ms_mean','iterations','dense_sweeps','push_edges_per_ns','source_edge_updates_per_s','error_flags','initial_solve_ms')})"
tail -2 gpurun_out/t58_c4_125.err | cut -c1-300
