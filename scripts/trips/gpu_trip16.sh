#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "carry" > gpurun_out/t16_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t16_tests.log; tail -3 gpurun_out/t16_tests.log
for env in "DPPR_CTAS_PER_SM=4" "DPPR_CTAS_PER_SM=2" "DPPR_CTAS_PER_SM=3" "DPPR_CTAS_PER_SM=2 DPPR_HUB_DEGREE=32" "DPPR_CTAS_PER_SM=2 DPPR_HUB_DEGREE=128" "DPPR_CTAS_PER_SM=1"; do
  echo "=== youtube $env"; env $env timeout 300 python scripts/probe.py --shape youtube --show 0 2>&1 | tail -5
done > gpurun_out/t16_probe.log 2>&1
grep -E "===|mean ms|per batch|push algo" gpurun_out/t16_probe.log
timeout 600 python bench.py > gpurun_out/t16_bench.json 2> gpurun_out/t16_bench.err; cat gpurun_out/t16_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('value','ms_per_step','p50_ms','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'], d['per_step'])"
