#!/bin/bash
# N-GPU: correctness worker of the host-buffer column split, then the bench line (HBM-resident value + end-to-end)
N=${1:-8}
mkdir -p gpurun_out/mg
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/mg_host_worker.py > gpurun_out/mg/worker_w$N.log 2>&1; echo "worker w$N rc=$?"; grep -v "^  File\|^    " gpurun_out/mg/worker_w$N.log | tail -$N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/mg/bench_n$N.log 2> gpurun_out/mg/bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/mg/bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d.get('e2e'))" || tail -20 gpurun_out/mg/bench_n$N.err
