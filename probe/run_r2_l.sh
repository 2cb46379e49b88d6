#!/bin/bash
out=gpurun_out/r2l; mkdir -p $out
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -8 $out/pytest.txt
timeout 300 python probe/perf_quick.py beta 2>&1 | tee $out/beta.txt
for v in trminb6 trminb8; do echo "== $v" | tee -a $out/geam_minb.txt; RTAT_B200_LIB=build/variants/librtat_b200_$v.so timeout 120 python probe/geam_transpose.py perf 2>&1 | grep -E "upc=0|sel=2 order=0" | tee -a $out/geam_minb.txt; done
echo "== shipped" | tee -a $out/geam_minb.txt; timeout 120 python probe/geam_transpose.py perf 2>&1 | grep -E "upc=0|sel=2 order=0" | tee -a $out/geam_minb.txt
timeout 120 python probe/sgemm_pair.py all 2>&1 | tail -10 | tee $out/sgemm.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.err
echo done
