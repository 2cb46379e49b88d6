#!/bin/bash
# N-GPU bench line only (HBM-resident value + end-to-end through the host-buffer column split), with the pipeline timeline
N=${1:-4}
mkdir -p gpurun_out/mg
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/mg/bench_n$N.log 2> gpurun_out/mg/bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/mg/bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])" || tail -20 gpurun_out/mg/bench_n$N.err
