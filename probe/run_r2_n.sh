#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -q -m gpu -k "ragged or peeled or config3 or stress or mid_size" ) > $out/pytest_peel.txt 2>&1; tail -5 $out/pytest_peel.txt
timeout 300 python probe/perf_quick.py c3pad 2>&1 | tee $out/c3.txt
echo done
