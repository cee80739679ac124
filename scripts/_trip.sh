mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "topk" > gpurun_out/u2_topk_tests.log 2>&1; tail -5 gpurun_out/u2_topk_tests.log
timeout 300 python bench.py --config 2 --steps 50 --warmup 5 --no-cpu > gpurun_out/u2_bench_c2.json 2>/dev/null
timeout 600 python bench.py --config 4 --sources 125 --steps 5 --warmup 3 --no-cpu > gpurun_out/u2_bench_c4_s125.json 2>/dev/null
timeout 900 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/u2_bench_c4.json 2>/dev/null
python - <<'PY'
import json
for f in ['u2_bench_c2','u2_bench_c4_s125','u2_bench_c4']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
