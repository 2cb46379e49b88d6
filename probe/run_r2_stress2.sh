#!/bin/bash
# Round 2, call B: which side of the ring protocol races?  fence before the release (arrive) vs fence after the acquire (wait),
# one CTA per SM, and the k-tile pattern inputs that give the direction of the stale data.
out=gpurun_out/r2b; mkdir -p $out
V=build/variants
C=tn_4096x1024x1024,pattern_tn_4096x1024x1024,nn_2048x2048x512,pattern_nn_2048x2048x512
for v in vec64 vec64_fa vec64_fw vec64_1cta split64park split64park_fa split64park_fw; do
  echo "== $v" | tee -a $out/stress.txt
  RTAT_B200_LIB=$V/librtat_b200_$v.so timeout 300 python probe/stress_dgemm.py --iters 3000 --tile 64 --traffic 0 --cases $C 2>&1 | tee -a $out/stress.txt
done
echo "== shipped, 20000 iterations, tile 64 / 128 / auto" | tee -a $out/stress.txt
for t in 64 128 0; do timeout 300 python probe/stress_dgemm.py --iters 20000 --tile $t --traffic 0 --cases $C 2>&1 | tee -a $out/stress.txt; done
echo done
