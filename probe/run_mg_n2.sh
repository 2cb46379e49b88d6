#!/bin/bash
# 2-GPU check of the multi-GPU host-buffer entry: correctness worker, then the bench with the pipeline timeline per flag transport
mkdir -p gpurun_out/mg
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/mg_host_worker.py > gpurun_out/mg/worker_w2.log 2>&1; echo "worker w2 rc=$?"; grep -v "^  File\|^    " gpurun_out/mg/worker_w2.log | tail -4
bash probe/run_mg_trace.sh 2 "pull push"
