#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python scripts/run_twitter.py --scale 1.0 --batches 10 --top-batches 3 --kinds rank1m,rank1k,top > gpurun_out/t22_tw_full.jsonl 2> gpurun_out/t22_tw_full.err; tail -3 gpurun_out/t22_tw_full.err; cat gpurun_out/t22_tw_full.jsonl
