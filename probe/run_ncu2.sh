#!/bin/bash
mkdir -p gpurun_out
python probe/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dgemm_ws|sgemm_tcgen05|geam_transpose_tma|geam_stream" -c 9 -o gpurun_out/r01_kernels -f python probe/ncu_target.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/r01_kernels.ncu-rep
