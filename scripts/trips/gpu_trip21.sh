#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/run_twitter.py --scale 0.05 --batches 5 > gpurun_out/t21_tw_small.jsonl 2> gpurun_out/t21_tw_small.err; tail -2 gpurun_out/t21_tw_small.err; cat gpurun_out/t21_tw_small.jsonl
timeout 1500 python scripts/run_twitter.py --scale 1.0 --batches 10 --kinds low > gpurun_out/t21_tw_full.jsonl 2> gpurun_out/t21_tw_full.err; tail -3 gpurun_out/t21_tw_full.err; cat gpurun_out/t21_tw_full.jsonl
