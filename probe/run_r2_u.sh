#!/bin/bash
out=gpurun_out/r2u; mkdir -p $out
echo "== fastsplit" | tee $out/sgemm.txt
RTAT_B200_LIB=build/variants/librtat_b200_fastsplit.so timeout 200 python probe/sgemm_pair.py all 2>&1 | tee -a $out/sgemm.txt
echo "== shipped" | tee -a $out/sgemm.txt
timeout 200 python probe/sgemm_pair.py all 2>&1 | tee -a $out/sgemm.txt
RTAT_B200_LIB=$PWD/build/variants/librtat_b200_fastsplit.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_l3.py -q -m gpu -k "sgemm or ssyrk or strsm" 2>&1 | tail -5 | tee $out/pytest_fast.txt
echo done
