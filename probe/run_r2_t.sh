#!/bin/bash
out=gpurun_out/r2t; mkdir -p $out
timeout 300 python probe/sgemm_c4_ops.py 2>&1 | tee $out/c4_ops.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err
echo done
