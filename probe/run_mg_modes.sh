#!/bin/bash
# e2e of the column-split bench under the three A-distribution modes (see bench.py RTAT_B200_MG_E2E)
N=${1:-2}
mkdir -p gpurun_out/mg
for mode in spin percall memop; do
  RTAT_B200_MG_E2E=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/mg/bench_n${N}_$mode.log 2> gpurun_out/mg/bench_n${N}_$mode.err
  tail -1 gpurun_out/mg/bench_n${N}_$mode.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$mode value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],1))" || tail -5 gpurun_out/mg/bench_n${N}_$mode.err
done
