#!/bin/bash
mkdir -p gpurun_out/geam
timeout 300 python -m pytest tests -x -q -m gpu -k "geam" 2>&1 | tail -5
timeout 120 python probe/perf_quick.py geam 2>&1 | tee gpurun_out/geam/perf.txt
