#!/bin/bash
mkdir -p gpurun_out
./build/atomics > gpurun_out/t3_atomics.log 2>&1; cat gpurun_out/t3_atomics.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_step_expand -s 30 -c 3 -o gpurun_out/prof_step python scripts/probe.py --shape youtube --mode 1 --batches 2 > gpurun_out/t3_ncu_step.log 2>&1
tail -3 gpurun_out/t3_ncu_step.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1a.csv python scripts/probe.py --shape youtube --batches 3 > gpurun_out/t3_ncu_list.log 2>&1
tail -3 gpurun_out/t3_ncu_list.log
ls -la gpurun_out
