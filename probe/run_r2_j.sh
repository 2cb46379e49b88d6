#!/bin/bash
out=gpurun_out/r2j; mkdir -p $out
./build/tma_coord_probe 2>&1 | tee $out/tma_coord.txt
timeout 300 python probe/geam_transpose.py all 2>&1 | tee $out/geam.txt
echo done
