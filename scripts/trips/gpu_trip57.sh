#!/bin/bash
# round-1 final measurements: full GPU suite, smoke, both bench arms, launch list + one full ncu capture of the bench kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t57_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t57_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/t57_bench.json 2> gpurun_out/t57_bench.err; echo "bench exit $?"; cat gpurun_out/t57_bench.json | cut -c1-1500
timeout 900 python bench.py --impl reference > gpurun_out/t57_bench_ref.json 2> gpurun_out/t57_bench_ref.err; echo "ref exit $?"; cat gpurun_out/t57_bench_ref.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/t57_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_persistent -s 5 -c 1 -o gpurun_out/prof_push_r01b python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/t57_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
