#!/bin/bash
mkdir -p gpurun_out
SEL="tests/test_gpu_dense.py::test_golden_with_forced_dense_iterations tests/test_gpu_dense.py::test_many_sources_lane_groups tests/test_gpu_dense.py::test_directed_out_lists_follow_the_window"
K="hub_expiry_directed or pl_undirected or lane_groups or coop"
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -q -x --timeout 550 -k "$K" > gpurun_out/t66_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/t66_$tool.log | tail -4
done
