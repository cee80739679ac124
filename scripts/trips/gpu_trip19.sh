#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 -x -k "golden_window and (hub_expiry or rmat_directed) and (levelsync or async)" > gpurun_out/t19_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/t19_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/t19_$tool.log | tail -4
done
