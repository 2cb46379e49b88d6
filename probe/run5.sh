#!/bin/bash
mkdir -p gpurun_out
EXP=host python probe/bisect_host.py seqNNN 1024 0 1 1 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; grep -v Warning gpurun_out/bench.err | head -20
for w in c1 c3; do timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline; timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --impl reference; done > gpurun_out/bench_c13.log 2>&1; cat gpurun_out/bench_c13.log | cut -c1-900
