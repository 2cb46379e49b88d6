#!/bin/bash
out=gpurun_out/r2k; mkdir -p $out
timeout 300 python probe/geam_transpose.py all 2>&1 | tee $out/geam.txt | tail -40
for v in chunk8 chunk16; do echo "== $v" | tee -a $out/sgemm.txt; RTAT_B200_LIB=build/variants/librtat_b200_$v.so timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee -a $out/sgemm.txt; done
python probe/ncu_c4_pair.py > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgemm_tcgen05_pair -s 1 -c 1 -o $out/c4_pair python probe/ncu_c4_pair.py > $out/ncu.log 2>&1
tail -3 $out/ncu.log
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -8 $out/pytest.txt
echo done
