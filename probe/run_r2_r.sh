#!/bin/bash
# Round 2, call R: SGEMM single 512-byte-row TMA box for M-contiguous A (against the four 128-byte boxes), then the evidence
# captures: ncu --set full of every hot kernel as shipped, and the launch list of the headline bench command
out=gpurun_out/r2r; mkdir -p $out
echo "== oldbox" | tee $out/sgemm.txt; RTAT_B200_LIB=build/variants/librtat_b200_oldbox.so timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee -a $out/sgemm.txt
echo "== shipped (one box)" | tee -a $out/sgemm.txt; timeout 200 python probe/sgemm_pair.py all 2>&1 | tail -12 | tee -a $out/sgemm.txt
( time timeout 900 python -m pytest tests -q -m gpu -k "sgemm or ssyrk or strsm or gemm_host or golden or dropin" ) > $out/pytest_s.txt 2>&1; tail -4 $out/pytest_s.txt
python probe/ncu_target.py > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dgemm_ws|sgemm_tcgen05|geam_transpose_reg|geam_stream" -c 12 -o $out/r02_kernels -f python probe/ncu_target.py > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
python probe/ncu_c4_pair.py > $out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgemm_tcgen05 -s 2 -c 2 -o $out/r02_sgemm -f python probe/ncu_c4_pair.py > $out/ncu_sgemm.log 2>&1
tail -2 $out/ncu_sgemm.log
echo done
