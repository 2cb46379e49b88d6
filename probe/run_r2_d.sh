#!/bin/bash
# Round 2, call D: (1) setmaxnreg with two CTAs per SM, idle producer warps EXITING, with and without the release fence;
# (2) the shipped build (fence + 16-byte fragments on both tile shapes) under the stress harness; (3) CTA-pair SGEMM.
out=gpurun_out/r2d; mkdir -p $out
V=build/variants
C=tn_4096x1024x1024,pattern_tn_4096x1024x1024,nn_2048x2048x512,pattern_nn_2048x2048x512
for v in split64_nofence split64_f2; do
  echo "== $v" | tee -a $out/stress.txt
  RTAT_B200_LIB=$V/librtat_b200_$v.so timeout 300 python probe/stress_dgemm.py --iters 3000 --tile 64 --traffic 0 --cases $C 2>&1 | cut -c1-220 | tee -a $out/stress.txt
done
echo "== shipped (fence + vec on 64 and 128)" | tee -a $out/stress.txt
timeout 600 python probe/stress_dgemm.py --iters 3000 --traffic 1 --pipeline 300 2>&1 | cut -c1-220 | tee -a $out/stress.txt
for t in 64 128; do timeout 300 python probe/stress_dgemm.py --iters 3000 --tile $t --traffic 0 2>&1 | cut -c1-220 | tee -a $out/stress.txt; done
echo "== sgemm pair" | tee $out/sgemm.txt
timeout 120 python probe/sgemm_pair.py check 2>&1 | tee -a $out/sgemm.txt
timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee -a $out/sgemm.txt
echo done
