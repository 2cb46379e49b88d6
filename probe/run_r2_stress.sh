#!/bin/bash
# Round 2, call A: iteration-count stress of the shipped DGEMM configurations and of the two parked variants
# (16-byte fragments on 64x64 tiles; setmaxnreg with two CTAs per SM, idle producer warps exiting vs parked).
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/clocks.txt
V=build/variants
echo "== shipped" | tee $out/stress.txt
timeout 600 python probe/stress_dgemm.py --iters 2000 --traffic 1 --pipeline 200 2>&1 | tee -a $out/stress.txt
for t in 64 128; do echo "== shipped, tile $t forced" | tee -a $out/stress.txt; timeout 300 python probe/stress_dgemm.py --iters 1000 --tile $t --traffic 1 2>&1 | tee -a $out/stress.txt; done
echo "== vec64 (16-byte fragments on 64x64 tiles)" | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_vec64.so timeout 600 python probe/stress_dgemm.py --iters 3000 --tile 64 --traffic 1 --pipeline 400 2>&1 | tee -a $out/stress.txt
echo "== vec64, no traffic" | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_vec64.so timeout 600 python probe/stress_dgemm.py --iters 3000 --tile 64 --traffic 0 2>&1 | tee -a $out/stress.txt
echo "== perf: shipped / vec64" | tee $out/perf.txt
timeout 300 python probe/perf_quick.py tiles 2>&1 | tee -a $out/perf.txt
RTAT_B200_LIB=$V/librtat_b200_vec64.so timeout 300 python probe/perf_quick.py tiles 2>&1 | grep -A8 "tile 64" | tee -a $out/perf.txt
timeout 200 python probe/perf_quick.py geam 2>&1 | tee -a $out/perf.txt
timeout 200 python probe/perf_quick.py sgemm 2>&1 | tee -a $out/perf.txt
echo "== split64park (setmaxnreg, 2 CTAs/SM, idle producer warps parked)" | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_split64park.so timeout 300 python probe/stress_dgemm.py --iters 1000 --tile 64 --traffic 1 --pipeline 100 2>&1 | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_split64park.so timeout 200 python probe/perf_quick.py tiles 2>&1 | grep -A8 "tile 64" | tee -a $out/perf.txt
echo "== vec64split64park" | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_vec64split64park.so timeout 300 python probe/stress_dgemm.py --iters 1000 --tile 64 --traffic 1 --pipeline 100 2>&1 | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_vec64split64park.so timeout 200 python probe/perf_quick.py tiles 2>&1 | grep -A8 "tile 64" | tee -a $out/perf.txt
echo "== split64 (idle producer warps exit)" | tee -a $out/stress.txt
RTAT_B200_LIB=$V/librtat_b200_split64.so timeout 120 python probe/stress_dgemm.py --iters 200 --tile 64 --traffic 0 2>&1 | tee -a $out/stress.txt
echo done
