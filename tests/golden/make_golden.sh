#!/bin/bash
# Regenerates tests/golden/reference_b200{,_l3}.{bin,json}: outputs of the UNMODIFIED reference executor
# (oracle/_ref/ref_driver = /root/reference/src compiled in place + oracle/ref_driver.cpp, cuBLAS 12.9)
# on seeded inputs.  Needs a GPU:   gpurun -- 'bash tests/golden/make_golden.sh'
# then copy gpurun_out/golden/* into tests/golden/.
set -e
mkdir -p gpurun_out/golden
oracle/_ref/ref_driver golden gpurun_out/golden/reference_b200.bin gpurun_out/golden/reference_b200.json
# SYRK / TRSM executors and raw calls (separate file so the GEMM-path goldens never change)
oracle/_ref/ref_driver golden_l3 gpurun_out/golden/reference_b200_l3.bin gpurun_out/golden/reference_b200_l3.json
ls -la gpurun_out/golden
