#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/t25_list.log 2>&1
tail -2 gpurun_out/t25_list.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_persistent -s 5 -c 1 -o gpurun_out/prof_push_r01 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/t25_full.log 2>&1
tail -2 gpurun_out/t25_full.log | cut -c1-300
timeout 900 ncu --set full --clock-control none -k regex:"repair_accumulate|radix_scatter|win_insert" -s 6 -c 6 -o gpurun_out/prof_twitter_r01 python scripts/run_twitter.py --scale 1.0 --batches 3 --kinds rank1m --check 0 > gpurun_out/t25_tw.log 2>&1
tail -2 gpurun_out/t25_tw.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
