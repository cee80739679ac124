#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/t37_dense.log
echo "dense tests exit $?"; tail -5 gpurun_out/t37_dense.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
for div in 0 8 4 2 16; do
  echo "=== DPPR_DENSE_DIV=$div youtube"; DPPR_DENSE_DIV=$div timeout 300 python scripts/probe.py --shape youtube --batches 30 --show 0 2>&1 | grep -E "mean ms|per batch"
done
