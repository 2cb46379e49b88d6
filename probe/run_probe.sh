#!/bin/bash
# First GPU probe: FP64 issue rates, cuBLAS numbers on the BASELINE configs, the reference's own
# run_tests driver on the same configs, and the cuBLAS kernel names (ncu launch list).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
probe/fp64_rates > gpurun_out/fp64_rates.txt 2>&1
probe/cublas_ref 5 > gpurun_out/cublas_ref.txt 2>&1
for c in c1 c2 c3 c4; do
  timeout 300 oracle/_ref/run_tests probe/inputs/$c.json > gpurun_out/ref_run_tests_$c.json 2> gpurun_out/ref_run_tests_$c.err
done
probe/cublas_ref 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/cublas_launches.csv probe/cublas_ref 1 > gpurun_out/ncu.log 2>&1
tail -40 gpurun_out/fp64_rates.txt; cat gpurun_out/cublas_ref.txt
