#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/t17_bench2.json 2> gpurun_out/t17_bench2.err
echo "exit $?"; tail -3 gpurun_out/t17_bench2.err; cat gpurun_out/t17_bench2.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('value','n_gpus','ms_per_step','estimates_gathered','gpu_launches')}, d['e2e'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/t17_ref2.json 2> gpurun_out/t17_ref2.err
echo "exit $?"; cat gpurun_out/t17_ref2.json | cut -c1-300
