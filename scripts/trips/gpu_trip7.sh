#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/t7_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t7_tests.log; tail -5 gpurun_out/t7_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t7_smoke.log 2>&1; tail -2 gpurun_out/t7_smoke.log
timeout 900 python bench.py > gpurun_out/t7_bench.json 2> gpurun_out/t7_bench.err; tail -3 gpurun_out/t7_bench.err; cat gpurun_out/t7_bench.json
timeout 900 python bench.py --impl reference --steps 20 > gpurun_out/t7_bench_ref.json 2> gpurun_out/t7_bench_ref.err; tail -3 gpurun_out/t7_bench_ref.err; cat gpurun_out/t7_bench_ref.json
