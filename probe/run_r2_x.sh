#!/bin/bash
out=gpurun_out/r2x; mkdir -p $out
echo "== column raster (before)" | tee $out/sgemm_big.txt
RTAT_B200_LIB=build/variants/librtat_b200_colraster.so timeout 200 python probe/sgemm_big.py 2>&1 | tee -a $out/sgemm_big.txt
echo "== grouped raster" | tee -a $out/sgemm_big.txt
timeout 200 python probe/sgemm_big.py 2>&1 | tee -a $out/sgemm_big.txt
timeout 200 python probe/sgemm_pair.py all 2>&1 | tail -10 | tee $out/sgemm.txt
( time timeout 900 python -m pytest tests -q -m gpu -k "sgemm or ssyrk or strsm or golden or dropin or l3" ) > $out/pytest_s.txt 2>&1; tail -4 $out/pytest_s.txt
echo done
