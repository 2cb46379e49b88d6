#!/bin/bash
mkdir -p gpurun_out
N=$1
export NCCL_DEBUG=WARN
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
print("N=$N value %.2f TF"%d["value"], "ms %.2f"%d["ms_per_step"], "per-gpu %.2f"%(d["value"]/$N), "roof", "%.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["value"], d["config"]["parallelism"], d["clocks"])
PY
grep -v "Warning\|W1020\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
