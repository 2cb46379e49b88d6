#!/bin/bash
mkdir -p gpurun_out/frag
timeout 120 python probe/perf_tma3d.py 2>&1 | grep "tma3d=1" | tee gpurun_out/frag/perf.txt
for i in 1 2 3 4; do timeout 60 python probe/debug_mg1.py 2>&1 | grep -c " ok"; timeout 60 python probe/debug_mg1.py 2>&1 | grep BAD | cut -c1-150; done
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/frag/pytest.txt
