#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/t51_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/t51_tests.log
for args in "--shape youtube --batches 50" "--shape orkut --scale 0.25 --batches 10" "--shape livejournal --scale 0.25 --batches 10"; do
  echo "=== $args"; DPPR_DENSE_MIN_EDGES=1000000 timeout 300 python scripts/probe.py $args --show 0 2>&1 | grep -E "mean ms"
done
DPPR_ITERLOG=1 timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 4 --top-batches 2 --kinds rank1k,top,rank1m --check 0 2>gpurun_out/t51_tw.err | tee gpurun_out/t51_tw.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('kind','push_ms_mean','step_ms_p50','iterations','dense_sweeps','push_edges_per_ns','error_flags')})"
grep "per-iteration" gpurun_out/t51_tw.err | cut -c1-400
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__inst_executed.sum --clock-control none -k regex:push_persistent -c 2 --csv --log-file gpurun_out/t51_ncu_twitter.csv python scripts/run_twitter.py --scale 1.0 --batches 1 --kinds rank1k --check 0 > gpurun_out/t51_ncu.out 2>&1
echo "ncu exit $?"; tail -25 gpurun_out/t51_ncu_twitter.csv | cut -c1-300
