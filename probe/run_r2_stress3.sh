#!/bin/bash
# Round 2, call C: which fence in front of the stage release is enough, and what each costs.
out=gpurun_out/r2c; mkdir -p $out
V=build/variants
C=tn_4096x1024x1024,pattern_tn_4096x1024x1024,nn_2048x2048x512
for v in vec64_f1 vec64_f2 vec64_f3; do
  echo "== $v" | tee -a $out/stress.txt
  RTAT_B200_LIB=$V/librtat_b200_$v.so timeout 300 python probe/stress_dgemm.py --iters 6000 --tile 64 --traffic 0 --cases $C 2>&1 | tee -a $out/stress.txt
done
for v in base base_f1 base_f2 base_f3 vec64_f1 vec64_f2 vec64_f3; do
  echo "== perf $v" | tee -a $out/perf.txt
  lib=$V/librtat_b200_$v.so; [ $v = base ] && lib=rtatblas_b200/lib/librtat_b200.so
  RTAT_B200_LIB=$lib timeout 300 python probe/perf_quick.py tiles 2>&1 | grep -E "^tile|C1|C2|C3|4096|2048" | tee -a $out/perf.txt
done
echo done
