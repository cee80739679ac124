#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_async -s 2 -c 1 -o gpurun_out/prof_async python scripts/probe.py --shape dblp --batches 3 --show 0 > gpurun_out/t12_ncu.log 2>&1
tail -3 gpurun_out/t12_ncu.log
