#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
python probe/ncu_c4_pair.py > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgemm_tcgen05 -s 2 -c 2 -o $out/c4_pair python probe/ncu_c4_pair.py > $out/ncu.log 2>&1
tail -3 $out/ncu.log
