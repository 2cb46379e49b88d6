#!/bin/bash
mkdir -p gpurun_out
python probe/ncu_l3_target.py > gpurun_out/plain_l3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dgemm_ws|sgemm_tcgen05|trsm_panel|trsm_diag" -c 12 -o gpurun_out/r01_l3_kernels -f python probe/ncu_l3_target.py > gpurun_out/ncu_l3.log 2>&1
tail -2 gpurun_out/ncu_l3.log; ls -la gpurun_out/r01_l3_kernels.ncu-rep
