#!/bin/bash
# first GPU trip: primitives + parity tests, short sanitizer pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/t1_gpu.txt 2>&1
nproc >> gpurun_out/t1_gpu.txt; lscpu | grep "Model name" >> gpurun_out/t1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_primitives.py -m gpu -x -q --timeout 120 > gpurun_out/t1_prims.log 2>&1
echo "prims exit $?" >> gpurun_out/t1_prims.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x > gpurun_out/t1_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/t1_parity.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -k "golden_window and batch_of_one and persistent" > gpurun_out/t1_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/t1_memcheck.log
tail -5 gpurun_out/t1_prims.log; tail -30 gpurun_out/t1_parity.log; tail -5 gpurun_out/t1_memcheck.log
