#!/bin/bash
mkdir -p gpurun_out
DPPR_ITERLOG=1 DPPR_PROBE_ITER=6 timeout 300 python scripts/probe.py --shape youtube --batches 10 --show 0 > gpurun_out/t36_iter6.log 2>&1
DPPR_ITERLOG=1 DPPR_PROBE_ITER=12 timeout 300 python scripts/probe.py --shape youtube --batches 10 --show 0 > gpurun_out/t36_iter12.log 2>&1
DPPR_ITERLOG=1 DPPR_PROBE_ITER=40 timeout 300 python scripts/probe.py --shape youtube --batches 10 --show 0 > gpurun_out/t36_iter40.log 2>&1
