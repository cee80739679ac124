#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/t29_bench.json 2> gpurun_out/t29_bench.err; tail -2 gpurun_out/t29_bench.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/t29_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','p50_ms','p95_ms','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['per_step'])
PY
timeout 900 python bench.py --impl reference --steps 20 > gpurun_out/t29_bench_ref.json 2> gpurun_out/t29_bench_ref.err; cut -c1-250 gpurun_out/t29_bench_ref.json
