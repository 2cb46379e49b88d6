#!/bin/bash
out=gpurun_out/r2e; mkdir -p $out
timeout 120 python probe/sgemm_pair.py check 2>&1 | tee $out/sgemm.txt
timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee -a $out/sgemm.txt
timeout 120 python probe/geam_copy.py 2>&1 | tee $out/geam_copy.txt
echo done
