#!/bin/bash
mkdir -p gpurun_out
python probe/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dgemm_ws -c 5 -o gpurun_out/r01_dgemm_ws -f python probe/ncu_target.py > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log; ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
