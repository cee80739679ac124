#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" DPPR_DENSE_DIV=128 timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 3 --kinds rank1k --check 0 2>gpurun_out/t44_$label.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$label', {k:d.get(k) for k in ('kind','push_ms_mean','iterations','push_edges_per_ns','error_flags')})"
}
run base4 DPPR_CTAS_PER_SM=4
run base5 DPPR_CTAS_PER_SM=5
run ipt2_4 DPPR_LIB=$PWD/dynamicppr_b200/lib/libdppr_ipt2.so DPPR_CTAS_PER_SM=4
run ipt2_6 DPPR_LIB=$PWD/dynamicppr_b200/lib/libdppr_ipt2.so DPPR_CTAS_PER_SM=6
run ipt2_8 DPPR_LIB=$PWD/dynamicppr_b200/lib/libdppr_ipt2.so DPPR_CTAS_PER_SM=8
