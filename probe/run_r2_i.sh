#!/bin/bash
# Round 2, call I: split-k TMA path of the DGEMM (odd leading dimensions), pair-SGEMM ring depth variants, whole GPU suite
out=gpurun_out/r2i; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "splitk or config3 or ragged" ) > $out/pytest_splitk.txt 2>&1; tail -5 $out/pytest_splitk.txt
timeout 300 python probe/perf_quick.py c3 2>&1 | tee $out/c3.txt
for v in pair_s4 pair_s5 pair_s7; do echo "== $v" | tee -a $out/sgemm.txt; RTAT_B200_LIB=build/variants/librtat_b200_$v.so timeout 120 python probe/sgemm_pair.py perf 2>&1 | grep "pair=1" | tee -a $out/sgemm.txt; done
echo "== shipped" | tee -a $out/sgemm.txt; timeout 120 python probe/sgemm_pair.py all 2>&1 | tail -12 | tee -a $out/sgemm.txt
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -8 $out/pytest.txt
echo done
