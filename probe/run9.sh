#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
cut -c1-1800 gpurun_out/bench2.log; grep -v "Warning\|W1020\|^\*\*\*" gpurun_out/bench2.err | tail -15
