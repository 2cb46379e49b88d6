#!/bin/bash
# N-GPU column-split bench with the host pipeline's per-stage timeline, shared-upload mode and per-call mode
N=${1:-2}
mkdir -p gpurun_out/mg
for mode in ${2:-pull push percall}; do
  RTAT_B200_HOST_TRACE=1 RTAT_B200_MG_E2E=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/mg/trace_n${N}_$mode.log 2> gpurun_out/mg/trace_n${N}_$mode.err
  tail -1 gpurun_out/mg/trace_n${N}_$mode.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$mode value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],1))" || tail -5 gpurun_out/mg/trace_n${N}_$mode.err
done
