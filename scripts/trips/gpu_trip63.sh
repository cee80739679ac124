#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 900 -x -k switching --durations=5 2>&1 | tail -12
