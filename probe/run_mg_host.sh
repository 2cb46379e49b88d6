#!/bin/bash
# Multi-GPU host-buffer entry: correctness worker at world 1 and world N, then the bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out/mg
timeout 300 python tests/mg_host_worker.py > gpurun_out/mg/worker_w1.log 2>&1; echo "worker w1 rc=$?"; tail -3 gpurun_out/mg/worker_w1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/mg_host_worker.py > gpurun_out/mg/worker_w$N.log 2>&1; echo "worker w$N rc=$?"; tail -5 gpurun_out/mg/worker_w$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/mg/bench_n$N.log 2> gpurun_out/mg/bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/mg/bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d.get('e2e'))" || tail -20 gpurun_out/mg/bench_n$N.err
