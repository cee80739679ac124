#!/bin/bash
# final verification of the round: full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t67_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t67_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/t67_bench.json 2> gpurun_out/t67_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/t67_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/t67_bench_ref.json 2> gpurun_out/t67_bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/t67_bench_ref.json
