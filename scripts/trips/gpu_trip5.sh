#!/bin/bash
mkdir -p gpurun_out
DPPR_ITERLOG=1 timeout 600 python scripts/probe.py --shape youtube --show 0 --batches 5 2>&1 | tail -12 > gpurun_out/t5_iterlog.log
DPPR_ITERLOG=1 DPPR_CTAS_PER_SM=1 timeout 600 python scripts/probe.py --shape youtube --show 0 --batches 5 2>&1 | tail -3 >> gpurun_out/t5_iterlog.log
cat gpurun_out/t5_iterlog.log
