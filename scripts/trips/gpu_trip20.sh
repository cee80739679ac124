#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 > gpurun_out/t20_full.log 2>&1; tail -3 gpurun_out/t20_full.log
for v in 0 1 2 3; do timeout 900 python scripts/run_config.py --config 3 --variant $v --check $((v==0)) 2>gpurun_out/t20_c3_$v.err | tee -a gpurun_out/t20_configs.jsonl; done
free -g | head -2
timeout 1500 python scripts/run_config.py --config 4 --sources 1 --batches 20 2>gpurun_out/t20_c4.err | tee -a gpurun_out/t20_configs.jsonl
